''' drone3d/raceline/base_raceline.py of the reference '''
from aircraft_trajectory_optimization_b200.raceline import RacelineConfig, GlobalRacelineConfig, \
    ParametricRacelineConfig, RacelineResults, BaseRaceline          # noqa: F401
