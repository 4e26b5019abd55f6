/*
 * inversus_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Scalar CPU restatement (plain C) of the reference rollout hot path of
 * Jason-Hoford/inversus-reinforcement-learning:
 *   inversus/core.py            (engine:  reset / move / shoot / wide shot / reload / bullets)
 *   inversus_rl/env_wrappers.py (wrapper: action decoding, scripted dummy, reward shaping,
 *                                done / timeout, 12-channel observation, vector runner)
 *   inversus_rl/training.py:140-151 (trainer-side auto-reset on done)
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may build, load or call this. The product path (inversus-reinforcement-learning_b200/)
 * never links or imports it.
 *
 * Parity status: PINNED. tests/golden/make_golden.py runs the live Python reference from
 * /root/reference with the draw shim below injected at its two RNG seams and commits the
 * resulting fixtures under tests/golden/; tests/test_oracle_golden.py replays them through
 * this file bit-for-bit, tests/test_oracle_reference_kat.py restates the reference's own
 * unit tests, and tests/test_oracle_vs_reference_live.py fuzzes it against the live
 * reference whenever /root/reference is present.
 *
 * Board size is a run-time parameter here (the reference's tests use 5x1 ... 15x10); the
 * CUDA product fixes 15x10.
 */
#ifndef INVERSUS_ORACLE_H
#define INVERSUS_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MAX_DIM 16
#define ORC_MAX_BULLETS 64

/* tile colours (game_types.py:8-11). P1 cannot walk on BLACK, P2 cannot walk on WHITE
 * (config.py:9,11). */
#define ORC_BLACK 0
#define ORC_WHITE 1

/* directions in the order the action ids and the observation channels use
 * (env_wrappers.py:24-37, :216-221): UP, RIGHT, DOWN, LEFT */
#define ORC_UP 0
#define ORC_RIGHT 1
#define ORC_DOWN 2
#define ORC_LEFT 3

/* RNG stream id used for reset draws (step draws use the pre-step step_count) */
#define ORC_STREAM_RESET 0xFFFFFFFFu
/* draw-table layout: per env per call, ORC_TABLE_STRIDE u32; dummy draws start at 0,
 * reset draws at ORC_TABLE_RESET_OFF */
#define ORC_TABLE_STRIDE 64
#define ORC_TABLE_RESET_OFF 16

typedef struct {
    int32_t x, y, dir, owner; /* owner 0 = P1, 1 = P2 */
} orc_bullet;

typedef struct {
    int32_t x, y, ammo, reload, alive;
} orc_player;

typedef struct {
    /* engine state (core.py:36-51) */
    int32_t width, height;
    uint8_t grid[ORC_MAX_DIM * ORC_MAX_DIM]; /* [y*width+x], ORC_BLACK/ORC_WHITE */
    orc_player p[2];
    int32_t n_bullets;
    orc_bullet bullets[ORC_MAX_BULLETS];
    /* wrapper state (env_wrappers.py:261-270) */
    int32_t step_count;
    int32_t prev_alive[2];
    double episode_return;
    /* configuration */
    int32_t max_episode_steps;
    int32_t difficulty; /* 0 easy, 1 hard */
    int32_t mode;       /* 0 dummy, 1 selfplay */
    /* draw source */
    uint32_t episode;    /* index of the running episode; bumped by every reset */
    uint32_t env_gid;    /* global env id, RNG key */
    uint64_t seed;
    const uint32_t *table; /* NULL -> Philox; else ORC_TABLE_STRIDE u32 for this env+call */
    /* bookkeeping for tests */
    int32_t draws_used;      /* draws consumed by the last dummy/reset call */
    int32_t bullet_overflow; /* sticky: >16 live bullets seen (device capacity) */
} orc_env;

typedef struct {
    double reward;
    int32_t done;
    int32_t landed_hit, got_hit, win, lose;
    int32_t episode_steps;
    double episode_return;
    int32_t a2; /* the P2 action id that was applied */
} orc_step_out;

/* ---- RNG ---- */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
uint32_t orc_draw_u32(uint64_t seed, uint32_t env_gid, uint32_t episode, uint32_t stream, uint32_t k);

/* ---- engine (core.py) ---- */
void orc_init(orc_env *e, int width, int height, int mode, int difficulty, int max_episode_steps,
              uint64_t seed, uint32_t env_gid);
void orc_engine_reset(orc_env *e);                 /* core.py:55-154, bumps episode */
int orc_try_move(orc_env *e, int pid, int dir);    /* core.py:249-296 */
int orc_spawn_bullet(orc_env *e, int pid, int dir);/* core.py:298-326 */
int orc_spawn_wide_shot(orc_env *e, int pid, int dir); /* core.py:328-381 */
void orc_reload_ammo(orc_env *e);                  /* core.py:383-397 */
void orc_update_bullets(orc_env *e);               /* core.py:399-475 */
int orc_is_round_over(const orc_env *e);           /* core.py:477-481 */
int orc_get_winner(const orc_env *e);              /* core.py:483-495; 0 none, 1 P1, 2 P2 */
void orc_apply_action(orc_env *e, int pid, int action_id);
void orc_step_players(orc_env *e, int a1, int a2); /* core.py:497-531 */

/* ---- wrapper (env_wrappers.py) ---- */
int orc_dummy_policy(orc_env *e);                  /* env_wrappers.py:69-170, returns action id */
void orc_build_obs(const orc_env *e, int viewer, float *grid12, float *extra4); /* :173-245 */
void orc_rl_reset(orc_env *e);                     /* env_wrappers.py:272-284 */
int orc_rl_step(orc_env *e, int a1, int a2, orc_step_out *out); /* :286-444; <0 on bad action */

/* ---- vector runner (env_wrappers.py:447-528 + training.py:140-151 auto-reset) ----
 * obs pointers may be NULL to skip a view. table, when non-NULL, holds n*ORC_TABLE_STRIDE u32.
 * Returns 0, or -1 if any action id is outside [0,12]. Fans out over envs with pthreads. */
int orc_batch_reset(orc_env *envs, int64_t n, const uint32_t *table, float *obs1, float *extra1,
                    float *obs2, float *extra2, int nthreads);
int orc_batch_step(orc_env *envs, int64_t n, const int8_t *a1, const int8_t *a2,
                   const uint32_t *table, int auto_reset, float *obs1, float *extra1, float *obs2,
                   float *extra2, float *reward, uint8_t *done, uint8_t *flags,
                   int32_t *episode_steps, double *episode_return, int nthreads);
int orc_max_threads(void);
int64_t orc_sizeof_env(void);

#ifdef __cplusplus
}
#endif
#endif
