''' drone3d/pytypes.py of the reference: configuration and state dataclasses '''
from aircraft_trajectory_optimization_b200.pytypes import *          # noqa: F401,F403
from aircraft_trajectory_optimization_b200.pytypes import PythonMsg, RacerConfig, PointConfig, DroneConfig, \
    RacerState, DroneState, PointState                                # noqa: F401
