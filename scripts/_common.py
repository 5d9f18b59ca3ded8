''' shared by the headless script mirrors: repo root on sys.path and the reference's result table '''
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def print_table(solvers, results):
    ''' scripts/fig_8.py:72-80 / race.py:78-86 / obstacles.py:46-55 of the reference '''
    print('Raceline       Times: \tLap \tIPOPT \tNLP \tTotal \tSetup')
    delim = 's'
    for (solver, result) in zip(solvers, results):
        print(f'{result.label:20s}' +
              f'\t{result.time:0.3f}' + delim +
              f'\t{result.ipopt_time:0.3f}' + delim +
              f'\t{result.feval_time:0.3f}' + delim +
              f'\t{result.solve_time:0.3f}' + delim +
              (f'\t{solver.setup_time:0.3f}' + delim if solver is not None else '') +
              f'\t{"feasible" if result.feasible else "NOT CONVERGED"}')
