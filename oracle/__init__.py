'''
CPU oracle for the raceline-NLP hot path.  TEST INFRASTRUCTURE ONLY.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs
may import or execute anything in this package; the product package
`aircraft_trajectory_optimization_b200` never does (and fails loudly without its CUDA library).

What it restates
----------------
The reference (`thomasfork/aircraft_trajectory_optimization`, all Python) only *describes* its
NLP as CasADi SX expressions and hands them to `ca.nlpsol('solver','ipopt',...)`
(drone3d/raceline/base_raceline.py:752-799).  The arithmetic of the hot path -- nlp_f, nlp_g,
nlp_grad_f, nlp_jac_g, nlp_hess_l -- lives in CasADi (third party, un-pinned in the reference's
setup.py:13, absent from this image, as is IPOPT).  This oracle therefore

  1. restates the reference's NLP construction literally (`ref_raceline.py`, `ref_models.py`,
     `ref_centerline.py`, `ref_discretization.py`, each function citing the file:line it follows)
     on a scalar expression DAG that applies CasADi-SX's construction-time folding rules,
  2. differentiates the fully unrolled graph the way CasADi does (source-transformation AD,
     structural sparsity by dependency propagation) and lays the results out in CasADi's
     conventions (jac_g: CCS ng x nw; hess_l: upper-triangular CCS of sigma*f + lam_g'g),
  3. evaluates the resulting flat instruction tapes with a small C interpreter (`sxvm.c`),
     which is also how CasADi executes SX functions (one interpreted scalar op per instruction).

The node store (`aircraft_trajectory_optimization_b200.symbolic.Graph`: hash-consed nodes,
folding rules, elementary partial derivatives) is shared with the product's build-time code
generator; everything above it (NLP construction, whole-graph AD, sparsity, CCS assembly,
evaluation) is independent of the product's per-interval kernels.  The shared elementary rules
are pinned separately in tests/ by finite differences and by sympy.

PARITY PIN STATUS: **parity unpinned against CasADi/IPOPT** -- the reference's only test
(tests/test_kinematics.py) pins the ODE right-hand sides (global == parametric trajectories),
which tests/test_oracle_kinematics.py reproduces with this restatement; no golden vector for
g / jac_g / hess_l / lap times exists in the reference and CasADi cannot be run here.
'''
