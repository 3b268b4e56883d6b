"""Roofline model of the fused Ant step kernel (used by bench.py and DESIGN.md; one formula, one place).

Algorithmic FLOPs per env sub-step, counted on the algorithm the kernel implements (block-arrow
elimination of the 14-dof tree, see csrc/hrl_ant.cuh), NOT on the redundant work some lanes
replicate.  C = contacts, L = active joint-limit rows, R = 3 C + L constraint rows:

    forward kinematics (quat -> R, 4 legs)                       430
    collision tests (13 spheres x ground/walls[/box])            200
    bias forces + leg blocks + Schur complement (4 legs)        2900
    6x6 Cholesky, L^-1, base/joint accelerations                 500
    integration (pose, quaternion exponential)                    80
    per row: Jacobian 30 + response M^-1 J^T 150                 180 R
    per row and solver iteration: 40                              40 R x iters

SURVEY.md 8(d) gave an a-priori 19.4 kFLOP/sub-step for a per-row ABA impulse response; the
elimination used here needs about half of that, and the lower number is the one reported.
"""

FIXED_FLOP_PER_SUBSTEP = 430 + 200 + 2900 + 500 + 80
ROW_FLOP = 180
ROW_ITER_FLOP = 40
TASK_FLOP_PER_STEP = 1500           # calc_state, 16-item sensor, reward/done
BYTES_PER_ENV_STEP = 605            # state R+W 232, items 128, counters 24, action 32, obs 184, rew+done 5
FP32_LANES_PER_SM = 128
N_SM = 148


def flop_per_env_step(contacts, limit_rows, substeps=4, iters=5):
    rows = 3.0 * contacts + limit_rows
    per_sub = FIXED_FLOP_PER_SUBSTEP + rows * (ROW_FLOP + ROW_ITER_FLOP * iters)
    return substeps * per_sub + TASK_FLOP_PER_STEP


def bytes_per_env_step(kind, obs_dim, act_dim):
    """Algorithmic HBM bytes per env-step: robot state 29 f32 read + written (232), the 16 item
    positions read (128, Gather envs only), counters / accumulators 24, the action read, the observation written,
    reward + done 5.  AntGather: 232 + 128 + 24 + 32 + 184 + 5 = 605."""
    items = 128 if kind in (0, 4) else 0   # HRL_ANT_GATHER, HRL_POINT_GATHER
    return 232 + items + 24 + 4 * act_dim + 4 * obs_dim + 5


def fp32_peak_tflops(sm_mhz):
    return N_SM * FP32_LANES_PER_SM * 2 * sm_mhz * 1e6 / 1e12
