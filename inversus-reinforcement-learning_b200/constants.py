"""Game and wrapper constants of the reference, mirrored once for the host side.

Values: inversus/config.py:7-17 and inversus_rl/env_wrappers.py:82-89. The device copies live in
csrc/inversus_kernels.cuh; tests/test_rng_spec.py checks the integer thresholds against the
reference's float comparisons.
"""
BOARD_W = 15
BOARD_H = 10
OBS_CHANNELS = 12
OBS_SHAPE = (OBS_CHANNELS, BOARD_H, BOARD_W)
OBS_ELEMS = OBS_CHANNELS * BOARD_H * BOARD_W
EXTRA_ELEMS = 4
NUM_ACTIONS = 13
MAX_AMMO = 6
RELOAD_TICKS_PER_AMMO = 30
WIDE_SHOT_AMMO_COST = 3
MAX_BULLETS = 16
PACKED_STATE_BYTES = 80

TABLE_STRIDE = 64
TABLE_RESET_OFF = 16
STREAM_RESET = 0xFFFFFFFF

# u = r / 2**32 with r a u32 draw; (u < p) <=> (r < THRESH)
THRESH_SHOOT_HARD = 858993460      # u < 0.2    (env_wrappers.py:88,96)
THRESH_RANDMOVE_HARD = 214748365   # u < 0.05   (env_wrappers.py:89,105)
THRESH_MOVE_EASY = 4294968         # not (u > 0.001)  (env_wrappers.py:82,123)

INFO_LANDED_HIT, INFO_GOT_HIT, INFO_WIN, INFO_LOSE = 1, 2, 4, 8
STATUS_INVALID_ACTION, STATUS_BULLET_OVERFLOW, STATUS_BAD_INDEX = 1, 2, 4

# algorithmic HBM bytes per env-step of the fused step kernel (DESIGN.md "roofline")
STATE_RW_BYTES = 2 * PACKED_STATE_BYTES
SMALL_OUT_BYTES = 16 + 4 + 1 + 1 + 4 + 8  # extra, reward, done, info, episode_steps, episode_return


def algorithmic_bytes_per_env_step(obs_elem_bytes: int = 4, selfplay: bool = False) -> int:
    views = 2 if selfplay else 1
    actions = 2 if selfplay else 1
    return (views * (OBS_ELEMS * obs_elem_bytes + 16) + (SMALL_OUT_BYTES - 16)
            + STATE_RW_BYTES + actions)
