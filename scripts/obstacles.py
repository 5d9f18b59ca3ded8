'''
obstacle-avoidance raceline inside an obstacle-free tube -- headless mirror of the reference's
scripts/obstacles.py:14-44 (waypoints, no gates, N = 100, collision radius 0.4).  The reference computes the
tube from a triangle mesh with trimesh (drone3d/obstacles/mesh_obstacle.py:38-145); trimesh is not installed
here, so the tube is injected through the `tube=` argument the reference's constructor already has
(base_raceline.py:1264-1271): the synthetic tube of SURVEY.md s8d (C3).
'''
import numpy as np

from _common import print_table
from aircraft_trajectory_optimization_b200.centerlines import SplineCenterline, SplineCenterlineConfig
from aircraft_trajectory_optimization_b200.collocation import get_collocation_coefficients
from aircraft_trajectory_optimization_b200.pytypes import DroneConfig
from aircraft_trajectory_optimization_b200.raceline import ParametricRacelineConfig, ParametricObstacleDroneRaceline, \
    ObstacleFreeTube


def synthetic_tube(line, N, K, collision_radius):
    tau = get_collocation_coefficients(K)[0]
    ds = (line.s_max() - line.s_min()) / N
    s = np.array([line.s_min() + ds * (n + tau[k]) for n in range(N) for k in range(K + 1)])
    ball_p = np.stack([s, 0.3 * np.sin(2 * np.pi * s / 12), 0.2 * np.cos(2 * np.pi * s / 12)], axis=1)
    return ObstacleFreeTube(ball_p, 0.9 * np.ones(len(s)), collision_radius)


def _main(verbose=False):
    x = np.array([-5, -2.75, -0.66, 2.95, 8.67, 9.2, 1.57, -2.39, -4.7, -2.39, 4.23, -2.66])
    y = np.array([4.5, -0.08, -1.36, 1.25, 6.69, -3.6, -6.43, -6, -6.43, -6.23, -0.66, 6.66])
    z = np.array([1.2, 2.815, 3.9, 2.815, 1.0, 1.0, 2.815, 3.9, 2.815, 1.0, 1.0, 1.0])
    config = SplineCenterlineConfig(x=np.array([x, y, z]))
    config.closed = True
    line = SplineCenterline(config)
    line.config.gate_s = None
    config = ParametricRacelineConfig(verbose=verbose, N=100)
    config.closed = line.config.closed
    drone_config = DroneConfig(global_r=True, use_quat=True, collision_radius=0.4)
    tube = synthetic_tube(line, config.N, config.K, drone_config.collision_radius)
    solver = ParametricObstacleDroneRaceline(line, config, drone_config, None, tube, generate_ws=True)
    raceline = solver.solve()
    print_table([solver, solver.ws_solver], [raceline, solver.ws_raceline])
    return solver, raceline


if __name__ == '__main__':
    import sys
    _main(verbose='-v' in sys.argv)
