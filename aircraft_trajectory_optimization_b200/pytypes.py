'''
Host-side data types kept API-compatible with the reference's `drone3d/pytypes.py`
(PythonMsg :11-46, state/vector types :48-330, RacerConfig :358, PointConfig :373,
DroneConfig :381-402, states :430-483).  Same class and field names, so code written against
the reference's configs and result states keeps working; the bodies are written fresh around
one generic "named vector" base instead of per-class boilerplate.
'''
from dataclasses import dataclass, field, fields
import copy

import numpy as np


@dataclass
class PythonMsg:
    ''' dataclass base that refuses to grow new attributes after construction '''

    def __setattr__(self, key, value):
        if not hasattr(self, key) and key not in getattr(self, '__dataclass_fields__', {}):
            raise TypeError(f'Not allowed to add new field "{key}" to class {self}')
        object.__setattr__(self, key, value)

    def copy(self):
        ''' deep copy '''
        return copy.deepcopy(self)

    def pprint(self, indent=0):
        ''' indented dump of all fields '''
        pad = ' ' * max(indent, 0)
        print(pad + type(self).__name__)
        for key, val in vars(self).items():
            if isinstance(val, PythonMsg):
                val.pprint(indent=max(indent, 0) + 2)
            else:
                print(f'{pad}  {key} : {val}')


@dataclass
class VectorizablePythonMsg(PythonMsg):
    ''' message whose float fields, in declaration order, form a vector '''

    def _vec_fields(self):
        return [f.name for f in fields(self)]

    def to_vec(self) -> np.ndarray:
        return np.array([getattr(self, k) for k in self._vec_fields()])

    def from_vec(self, vec) -> None:
        names = self._vec_fields()
        vec = list(vec)
        if len(vec) != len(names):
            raise ValueError(f'expected {len(names)} entries, got {len(vec)}')
        for k, v in zip(names, vec):
            object.__setattr__(self, k, v)


@dataclass
class Position(VectorizablePythonMsg):
    ''' global-frame position '''
    xi: float = 0
    xj: float = 0
    xk: float = 0


@dataclass
class BodyPosition(VectorizablePythonMsg):
    ''' body-frame position '''
    x1: float = 0
    x2: float = 0
    x3: float = 0


@dataclass
class BodyLinearVelocity(VectorizablePythonMsg):
    ''' body-frame linear velocity '''
    v1: float = 0
    v2: float = 0
    v3: float = 0

    def mag(self):
        return float(np.linalg.norm(self.to_vec()))

    def signed_mag(self):
        return self.mag() * np.sign(self.v1)


@dataclass
class BodyAngularVelocity(VectorizablePythonMsg):
    ''' body-frame angular velocity '''
    w1: float = 0
    w2: float = 0
    w3: float = 0


@dataclass
class BodyLinearAcceleration(VectorizablePythonMsg):
    ''' body-frame linear acceleration '''
    a1: float = 0
    a2: float = 0
    a3: float = 0


@dataclass
class BodyAngularAcceleration(VectorizablePythonMsg):
    ''' body-frame angular acceleration '''
    a1: float = 0
    a2: float = 0
    a3: float = 0


def quat_to_matrix(q) -> np.ndarray:
    ''' rotation matrix of a unit quaternion in the reference's [qi, qj, qk, qr] order '''
    qi, qj, qk, qr = q
    return np.array([
        [1 - 2 * (qj * qj + qk * qk), 2 * (qi * qj - qk * qr), 2 * (qi * qk + qj * qr)],
        [2 * (qi * qj + qk * qr), 1 - 2 * (qi * qi + qk * qk), 2 * (qj * qk - qi * qr)],
        [2 * (qi * qk - qj * qr), 2 * (qj * qk + qi * qr), 1 - 2 * (qi * qi + qj * qj)]])


def matrix_to_quat(R) -> np.ndarray:
    ''' [qi, qj, qk, qr] of a rotation matrix (scipy convention is the same scalar-last order) '''
    from scipy.spatial.transform import Rotation
    return Rotation.from_matrix(np.asarray(R, dtype=float)).as_quat()


def ypr_to_matrix(a, b, c) -> np.ndarray:
    ''' Ra(yaw a) @ Rb(pitch b) @ Rc(roll c) '''
    ca_, sa = np.cos(a), np.sin(a)
    cb, sb = np.cos(b), np.sin(b)
    cc, sc = np.cos(c), np.sin(c)
    Ra = np.array([[ca_, -sa, 0], [sa, ca_, 0], [0, 0, 1]])
    Rb = np.array([[cb, 0, sb], [0, 1, 0], [-sb, 0, cb]])
    Rc = np.array([[1, 0, 0], [0, cc, -sc], [0, sc, cc]])
    return Ra @ Rb @ Rc


@dataclass
class OrientationQuaternion(VectorizablePythonMsg):
    ''' Euler symmetric parameters, scalar last '''
    qi: float = 0
    qj: float = 0
    qk: float = 0
    qr: float = 1

    def R(self):
        return quat_to_matrix(self.to_vec())

    def Rinv(self):
        return self.R().T

    def e1(self):
        return self.R()[:, 0]

    def e2(self):
        return self.R()[:, 1]

    def e3(self):
        return self.R()[:, 2]

    def norm(self):
        return float(np.linalg.norm(self.to_vec()))

    def normalize(self):
        self.from_vec(self.to_vec() / self.norm())

    def from_yaw(self, yaw):
        self.from_vec([0, 0, np.sin(yaw / 2), np.cos(yaw / 2)])

    def to_yaw(self):
        return 2 * np.arctan2(self.qk, self.qr)

    def from_mat(self, R):
        self.from_vec(matrix_to_quat(R))

    def qdot(self, w: BodyAngularVelocity) -> 'OrientationQuaternion':
        ''' quaternion rate from body rates (same kinematics as rotations.py:67-80) '''
        qi, qj, qk, qr = self.to_vec()
        M = 0.5 * np.array([[qr, -qk, qj], [qk, qr, -qi], [-qj, qi, qr], [-qi, -qj, -qk]])
        out = OrientationQuaternion()
        out.from_vec(M @ w.to_vec())
        return out


@dataclass
class ParametricPosition(VectorizablePythonMsg):
    ''' curvilinear position (s, y, n) '''
    s: float = 0.
    y: float = 0.
    n: float = 0.


@dataclass
class Orientation(VectorizablePythonMsg):
    ''' some 3D orientation measure '''


@dataclass
class GlobalOrientation(Orientation):
    ''' orientation relative to the global frame '''


@dataclass
class RelativeOrientation(Orientation):
    ''' orientation relative to the centerline frame '''


@dataclass
class EulerAngles(VectorizablePythonMsg):
    ''' yaw a, pitch b, roll c '''
    a: float = 0.
    b: float = 0.
    c: float = 0.

    def R(self):
        return ypr_to_matrix(self.a, self.b, self.c)

    def from_mat(self, R):
        from scipy.spatial.transform import Rotation
        cba = Rotation.from_matrix(np.asarray(R, dtype=float)).as_euler('xyz', degrees=False)
        self.from_vec(cba[::-1])


@dataclass
class GlobalEulerAngles(EulerAngles, GlobalOrientation):
    ''' global Euler angles '''


@dataclass
class RelativeEulerAngles(EulerAngles, RelativeOrientation):
    ''' relative Euler angles '''


@dataclass
class GlobalQuaternion(OrientationQuaternion, GlobalOrientation):
    ''' global orientation quaternion '''


@dataclass
class RelativeQuaternion(OrientationQuaternion, RelativeOrientation):
    ''' relative orientation quaternion '''


@dataclass
class RacerConfig(PythonMsg):
    ''' generic vehicle config (pytypes.py:358-371) '''
    dt: float = 0.1
    m: float = 1.0
    g: float = 9.81
    # linear drag F = -b * v
    b1: float = 0
    b2: float = 0
    b3: float = 0
    global_r: bool = False
    collision_radius: float = 0.3


@dataclass
class PointConfig(RacerConfig):
    ''' point-mass config (pytypes.py:373-379) '''
    T_max: float = 32.4
    T_min: float = -32.4
    dT_max: float = 350
    dT_min: float = -350


@dataclass
class DroneConfig(RacerConfig):
    ''' quadrotor config (pytypes.py:381-402) '''
    I1: float = 1.0e-3
    I2: float = 1.0e-3
    I3: float = 1.7e-3
    l: float = 0.15
    k: float = 0.05
    T_max: float = 8.1
    T_min: float = 0.2
    dT_max: float = 20
    dT_min: float = -20
    # angular drag K = -bw * w
    bw1: float = 1e-4
    bw2: float = 1e-4
    bw3: float = 1e-4
    w_max: float = 10
    w_min: float = -10
    use_quat: bool = False


@dataclass
class DroneActuation(VectorizablePythonMsg):
    ''' four rotor thrusts '''
    u1: float = 0.
    u2: float = 0.
    u3: float = 0.
    u4: float = 0.


@dataclass
class PointActuation(VectorizablePythonMsg):
    ''' body-frame thrust vector '''
    u1: float = 0.
    u2: float = 0.
    u3: float = 0.


@dataclass
class RacerState(PythonMsg):
    ''' dynamic state of any racer '''
    t: float = 0.
    x: Position = field(default_factory=Position)
    q: OrientationQuaternion = field(default_factory=OrientationQuaternion)
    v: BodyLinearVelocity = field(default_factory=BodyLinearVelocity)
    w: BodyAngularVelocity = field(default_factory=BodyAngularVelocity)
    p: ParametricPosition = field(default_factory=ParametricPosition)
    # collision-avoidance distance
    d: float = 0


@dataclass
class DroneState(RacerState):
    ''' dynamic state of a drone '''
    r: Orientation = field(default_factory=RelativeQuaternion)
    u: DroneActuation = field(default_factory=DroneActuation)
    du: DroneActuation = field(default_factory=DroneActuation)


@dataclass
class PointState(RacerState):
    ''' dynamic state of a point mass '''
    u: PointActuation = field(default_factory=PointActuation)
    du: PointActuation = field(default_factory=PointActuation)
