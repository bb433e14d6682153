"""cts_b200 -- B200-native (sm_100a) per-step state-update path of CliMA/ClimaTimeSteppers.jl.

Host mirror of the reference's public API for that path (``ClimaODEFunction``, ``IMEXAlgorithm`` /
``ExplicitAlgorithm`` with the named tableaux, ``LSRK54CarpenterKennedy``, ``NewtonsMethod``, ``init`` /
``step!`` / ``solve!``) on top of the C-ABI library ``libcts_sm100a.so`` (``include/cts.h``).
The directory is called ``climatimesteppers.jl_b200``; import it as ``cts_b200`` (see ``cts_b200.py``).
"""
from . import _lib
from ._lib import CTSError, LIB_PATH
from .algorithms import (SSP, ConstantForcing, ConvergenceChecker, LineSearch, MaximumError, MaximumErrorReduction,
                         MaximumRelativeError, MinimumRateOfConvergence, MultipleConditions, EisenstatWalkerForcing, ExplicitAlgorithm, ForwardDiffJVP,
                         ForwardDiffStepSize1, ForwardDiffStepSize2, ForwardDiffStepSize3, IMEXAlgorithm,
                         KrylovMethod, LowStorageRungeKutta2N, LSRK54CarpenterKennedy, LSRK144NiegemannDiehlBusch,
                         LSRKEulerMethod, NewtonsMethod, RosenbrockAlgorithm, TimeSteppingAlgorithm, Unconstrained)
from .device import B200Vector, Context, axpy_n
from .functions import (ClimaODEFunction, EndOfStage, EndOfStep, IncrementingODEFunction, NewNewtonIteration,
                        NewNewtonSolve, NewTimeStep, ODEFunction, ODEProblem, ODESolution, Returns_nothing,
                        UpdateEvery, UpdateEveryDt, UpdateEveryN, WithDSS, has_T_exp, has_T_lim)
from .integrators import (CallbackSet, DiscreteCallback, EveryXSimulationSteps, EveryXSimulationTime,
                          EveryXWallTimeSeconds, HostSaver, add_saveat_, add_tstop_, get_dt, init, reinit_, set_dt_, solve,
                          solve_, step_)
from .operators import AdvectionOperator, ColumnOperator, ColumnTridiag
from .parallel import SlabPartition, attach_communicator
from .solvers import (IMEXCache, LSRKCache, NewtonCache, RosenbrockCache, allocate_cache, init_cache, solve_newton_,
                      step_u)
from .tableaus import *  # noqa: F401,F403  (named tableaux: ARS343, SSP333, ...)
from .tableaus import IMEX_NAMES, EXPLICIT_NAMES, IMEXTableau, RosenbrockTableau, SSPKnoth, is_fsal, tableau
