'''
Mesh obstacles: the numpy checker against the analytic distance of a box (CPU), the CUDA
brute-force kernel against the checker on a synthetic mesh and -- when the reference's arena mesh can be found -- on the
real one, and the obstacle-free tube built from it (GPU).
'''
import os

import numpy as np
import pytest

from oracle.ref_mesh import box_triangles, box_sdf, signed_distance

BOXES = [((-1.0, -0.5, 0.0), (0.5, 0.7, 1.2)), ((1.5, 1.0, -0.3), (2.5, 2.0, 0.9)), ((-3.0, 2.0, 0.5), (-2.0, 2.4, 2.5))]


def _mesh():
    return np.concatenate([box_triangles(lo, hi) for lo, hi in BOXES])


def _points(n, seed=0):
    return np.random.default_rng(seed).uniform([-4, -2, -1], [4, 4, 3], (n, 3))


def test_checker_matches_analytic_box_distance():
    X = _points(400)
    d, cp = signed_distance(_mesh(), X)
    ref = np.min([box_sdf(lo, hi, X) for lo, hi in BOXES], axis=0)        # disjoint boxes: union = min
    assert np.abs(d - ref).max() <= 1e-12
    assert np.abs(np.linalg.norm(X - cp, axis=1) - np.abs(d)).max() <= 1e-12
    assert (d < 0).sum() > 5 and (d > 0).sum() > 100                       # both signs are exercised


def test_obj_loader_and_shim_paths(tmp_path):
    from aircraft_trajectory_optimization_b200.obstacles import load_obj_triangles
    p = tmp_path / 'quad.obj'
    p.write_text('v 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nvn 0 0 1\nf 1//1 2//1 3//1 4//1\nf -4 -3 -2\n')
    T = load_obj_triangles(str(p))
    assert T.shape == (3, 3, 3) and np.allclose(T[1], [[0, 0, 0], [1, 1, 0], [0, 1, 0]])
    # the reference's module paths (scripts/*.py import these names)
    import importlib
    for mod, names in (('drone3d.utils.solve_util', ['solve_util']),
                       ('drone3d.centerlines.spline_centerline', ['SplineCenterline', 'SplineCenterlineConfig']),
                       ('drone3d.centerlines.base_centerline', ['GateShape']),
                       ('drone3d.pytypes', ['DroneConfig', 'PointConfig']),
                       ('drone3d.raceline.base_raceline', ['GlobalRacelineConfig', 'ParametricRacelineConfig']),
                       ('drone3d.raceline.drone_raceline', ['GlobalDroneRaceline', 'ParametricObstacleDroneRaceline']),
                       ('drone3d.obstacles.mesh_obstacle', ['MeshObstacle']),
                       ('drone3d.utils.load_utils', ['get_assets_file']),
                       ('drone3d.utils.cpc_utils', ['package_cpc_data_as_raceline']),
                       ('drone3d.visualization.drone_raceline_fig', ['DroneRacelineWindow'])):
        m = importlib.import_module(mod)
        assert all(hasattr(m, nme) for nme in names), mod


@pytest.mark.gpu
def test_cuda_sdf_matches_checker(built_library):
    from aircraft_trajectory_optimization_b200.obstacles import MeshObstacle
    T = _mesh()
    mesh = MeshObstacle(triangles=T)
    X = _points(3000, seed=1)
    d = mesh.signed_distance(X)
    cp = mesh.closest_point(X)
    dr, cpr = signed_distance(T, X)
    assert np.abs(d - dr).max() <= 1e-12
    assert np.abs(cp - cpr).max() <= 1e-12
    assert np.abs(d - np.min([box_sdf(lo, hi, X) for lo, hi in BOXES], axis=0)).max() <= 1e-12
    assert mesh.check_for_collisions(X[d > 0.5], 0.3) and not mesh.check_for_collisions(X, 0.3)


@pytest.mark.gpu
def test_tube_on_synthetic_mesh_and_solve(built_library):
    ''' compute_plannning_tube -> tube discs free of the mesh, usable by the obstacle raceline builders '''
    from cases import make_line
    from aircraft_trajectory_optimization_b200.obstacles import MeshObstacle
    line = make_line('obs')
    # two slabs well away from the centerline plus one near it: the tube must shrink / shift there
    xc = np.asarray(line.p2xc(3.0)).ravel()
    ey = np.asarray(line.p2ey(3.0)).ravel()
    near = xc + 0.8 * ey
    T = np.concatenate([box_triangles(near - 0.25, near + 0.25), box_triangles((20, 20, 20), (21, 21, 21))])
    mesh = MeshObstacle(triangles=T)
    s = np.linspace(line.s_min(), line.s_max(), 200, endpoint=False)
    tube = mesh.compute_plannning_tube(line, s, collision_r=0.2)
    assert tube.ball_p.shape == (200, 3) and (tube.ball_r > 0).all()
    # every sphere is empty: its radius equals the distance of its centre to the mesh
    d = mesh.signed_distance(tube.ball_center)
    assert np.abs(d - tube.ball_r).max() <= 1e-12
    k = int(np.argmin(np.abs(s - 3.0)))
    assert tube.ball_r[k] < tube.ball_r.max() and tube.ball_p[k, 1] <= 0.0      # pushed away from the box (towards -ey)


@pytest.mark.gpu
def test_real_arena_mesh_if_available(built_library):
    from aircraft_trajectory_optimization_b200.obstacles import MeshObstacle, _find_asset, load_obj_triangles
    try:
        path = _find_asset('arena_track_obstacles_multistory.obj')
    except FileNotFoundError:
        pytest.skip('the arena mesh is an asset of the reference repository (not shipped)')
    T = load_obj_triangles(path)
    mesh = MeshObstacle(triangles=T)
    X = np.random.default_rng(2).uniform(T.reshape(-1, 3).min(0), T.reshape(-1, 3).max(0), (64, 3))
    d = mesh.signed_distance(X)
    dr, _ = signed_distance(T, X)
    assert np.abs(np.abs(d) - np.abs(dr)).max() <= 1e-11 and (np.sign(d) == np.sign(dr)).all()
