''' drone3d/raceline/drone_raceline.py of the reference '''
from aircraft_trajectory_optimization_b200.raceline import GlobalDroneRaceline, ParametricDroneRaceline, \
    ParametricObstacleDroneRaceline                                   # noqa: F401
