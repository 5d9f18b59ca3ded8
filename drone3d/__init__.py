'''
Import-path shim: the module layout of the reference package (thomasfork/aircraft_trajectory_optimization, `drone3d/`),
re-exporting the B200 implementation in `aircraft_trajectory_optimization_b200`, so that the reference's
`scripts/*.py` (`from drone3d.utils.solve_util import solve_util`, ...) run unmodified and headless on the GPU path.
Only names are provided here; every class and function lives in the package it is imported from.
'''
