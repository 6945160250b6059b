"""B200-native batched contact-implicit interior-point path (drop-in for the MPC inner loop of
dojo-sim/ContactImplicitMPC.jl).  The directory name carries a dot, so import it through the
repo-root shim:  `import cimpc_b200`."""
from .capi import LIB_PATH, SYMBOLS, CimpcError, load_library  # noqa: F401
from .solver import (ImplicitTrajectory, InteriorPointOptions, Newton, NewtonOptions,  # noqa: F401
                     Simulator, implicit_dynamics, simulator_options)
from .sharding import (gather_rollout_results, gather_rollouts_capi, init_comm, shard_rollouts,  # noqa: F401,E402
                       sum_statistics)
from .rollout import GroupedRollouts, MonteCarloRollouts, ReferenceWindow, quadruped_initial_configurations  # noqa: F401,E402
from .trajectory import ContactTraj, JLD2File, load_gait, load_traj, repeat_ref_traj, save_gait, save_traj, tracking_error  # noqa: F401,E402
from .disturbances import (Disturbances, EmptyDisturbances, ImpulseDisturbance, OpenLoopDisturbance,  # noqa: F401,E402
                           RandomDisturbance)
