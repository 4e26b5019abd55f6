/*
 * inversus_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (see inversus_oracle.h).
 *
 * Plain-C scalar restatement of the reference's rollout hot path. Every function cites the
 * reference file:line it follows (paths relative to the reference root). The structure is
 * deliberately the naive one (byte-per-tile grid, bullet list, per-tile grouping) so that it
 * shares no implementation idea with the bit-packed CUDA product it is used to check.
 *
 * Parity status: PINNED against the live Python reference (tests/golden/make_golden.py,
 * tests/test_oracle_golden.py, tests/test_oracle_vs_reference_live.py).
 */
#include "inversus_oracle.h"

#include <pthread.h>
#include <string.h>
#include <unistd.h>

/* config.py:7-17 */
#define MAX_AMMO 6
#define RELOAD_TICKS_PER_AMMO 30
#define WIDE_SHOT_AMMO_COST 3
#define DEFAULT_START_X 1
#define DEFAULT_START_Y 1

static const int DIR_DX[4] = {0, 1, 0, -1};
static const int DIR_DY[4] = {-1, 0, 1, 0};
/* player colour = the colour the player can NOT stand on (game_types.py:58, config.py:9,11) */
static const int PLAYER_COLOR[2] = {ORC_BLACK, ORC_WHITE};

/* ------------------------------------------------------------------ RNG */

/* Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11).
 * Not part of the reference: it is the counter-based stream that replaces the reference's
 * Mersenne Twister at its two injection seams (core.py:41 env.rng, env_wrappers.py:5 random). */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* draw k of (env, episode, stream): counter = (env_gid, episode, stream, k/4), key = seed */
uint32_t orc_draw_u32(uint64_t seed, uint32_t env_gid, uint32_t episode, uint32_t stream, uint32_t k)
{
    uint32_t ctr[4] = {env_gid, episode, stream, k >> 2};
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint32_t out[4];
    orc_philox4x32_10(ctr, key, out);
    return out[k & 3];
}

static uint32_t next_draw(orc_env *e, int is_reset)
{
    uint32_t k = (uint32_t)e->draws_used++;
    if (e->table)
        return e->table[(is_reset ? ORC_TABLE_RESET_OFF : 0) + k];
    return orc_draw_u32(e->seed, e->env_gid, e->episode,
                        is_reset ? ORC_STREAM_RESET : (uint32_t)e->step_count, k);
}

/* random.random() stand-in: u = r / 2^32, exact in binary64 */
static double draw_random(orc_env *e) { return (double)next_draw(e, 0) * (1.0 / 4294967296.0); }

/* random._randbelow(n) stand-in: floor(r * n / 2^32) */
static int draw_below(orc_env *e, int n, int is_reset)
{
    return (int)(((uint64_t)next_draw(e, is_reset) * (uint64_t)n) >> 32);
}

/* random.randint(a, b); b < a returns a (lets the reference's height=1 tests construct) */
static int draw_randint(orc_env *e, int a, int b, int is_reset)
{
    int n = b - a + 1;
    if (n <= 0) { (void)next_draw(e, is_reset); return a; }
    return a + draw_below(e, n, is_reset);
}

/* random.shuffle(x): CPython Lib/random.py -- for i in reversed(range(1, len(x))):
 * j = randbelow(i + 1); x[i], x[j] = x[j], x[i] */
static void draw_shuffle(orc_env *e, int *x, int len)
{
    for (int i = len - 1; i >= 1; --i) {
        int j = draw_below(e, i + 1, 0);
        int t = x[i]; x[i] = x[j]; x[j] = t;
    }
}

/* ------------------------------------------------------------------ engine: core.py */

static int in_bounds(const orc_env *e, int x, int y) /* core.py:222-224 */
{
    return 0 <= x && x < e->width && 0 <= y && y < e->height;
}
static int get_tile(const orc_env *e, int x, int y) { return e->grid[y * e->width + x]; }
static void set_tile(orc_env *e, int x, int y, int c) { e->grid[y * e->width + x] = (uint8_t)c; }

/* core.py:238-247 */
static int walkable_for(const orc_env *e, int x, int y, int color)
{
    if (!in_bounds(e, x, y)) return 0;
    return get_tile(e, x, y) != color;
}

static void paint_plus(orc_env *e, int cx, int cy, int color)
{
    /* centre, right, left, down, up -- each clipped to the board (core.py:99-108) */
    static const int ox[5] = {0, 1, -1, 0, 0};
    static const int oy[5] = {0, 0, 0, 1, -1};
    for (int i = 0; i < 5; ++i) {
        int x = cx + ox[i], y = cy + oy[i];
        if (in_bounds(e, x, y)) set_tile(e, x, y, color);
    }
}

void orc_init(orc_env *e, int width, int height, int mode, int difficulty, int max_episode_steps,
              uint64_t seed, uint32_t env_gid)
{
    memset(e, 0, sizeof(*e));
    e->width = width;
    e->height = height;
    e->mode = mode;
    e->difficulty = difficulty;
    e->max_episode_steps = max_episode_steps;
    e->seed = seed;
    e->env_gid = env_gid;
    e->episode = 0xFFFFFFFFu; /* the first reset starts episode 0 */
    e->prev_alive[0] = e->prev_alive[1] = 1;
}

/* core.py:55-154 (InversusEnv.reset) */
void orc_engine_reset(orc_env *e)
{
    e->episode += 1u;
    e->draws_used = 0;
    /* config.py:20-56 make_initial_grid: all P1 colour, WHITE plus at the legacy start (1,1) */
    memset(e->grid, ORC_BLACK, sizeof(e->grid));
    if (in_bounds(e, DEFAULT_START_X, DEFAULT_START_Y))
        paint_plus(e, DEFAULT_START_X, DEFAULT_START_Y, ORC_WHITE);

    /* core.py:69-70 */
    int p1x = draw_randint(e, 1, e->width - 2, 1);
    int p1y = draw_randint(e, 1, e->height - 2, 1);
    /* core.py:85-90: up to 20 tries, the last draw is kept even if it is too close */
    int p2x = 0, p2y = 0;
    for (int t = 0; t < 20; ++t) {
        p2x = draw_randint(e, 1, e->width - 2, 1);
        p2y = draw_randint(e, 1, e->height - 2, 1);
        int dx = p2x - p1x, dy = p2y - p1y;
        int dist = (dx < 0 ? -dx : dx) + (dy < 0 ? -dy : dy);
        if (dist > 4) break;
    }
    paint_plus(e, p2x, p2y, ORC_WHITE); /* core.py:96-108 */
    paint_plus(e, p1x, p1y, ORC_WHITE); /* core.py:112-121 */
    paint_plus(e, p2x, p2y, ORC_BLACK); /* core.py:136-146 (wins every overlap) */

    e->p[0].x = p1x; e->p[0].y = p1y; e->p[0].ammo = MAX_AMMO; e->p[0].reload = 0; e->p[0].alive = 1;
    e->p[1].x = p2x; e->p[1].y = p2y; e->p[1].ammo = MAX_AMMO; e->p[1].reload = 0; e->p[1].alive = 1;
    e->n_bullets = 0; /* core.py:154 */
}

/* core.py:249-296 */
int orc_try_move(orc_env *e, int pid, int dir)
{
    orc_player *pl = &e->p[pid];
    if (!pl->alive) return 0;
    int nx = pl->x + DIR_DX[dir], ny = pl->y + DIR_DY[dir];
    if (!in_bounds(e, nx, ny)) return 0;
    if (!walkable_for(e, nx, ny, PLAYER_COLOR[pid])) return 0;
    pl->x = nx; pl->y = ny;
    return 1;
}

static void append_bullet(orc_env *e, int x, int y, int dir, int owner)
{
    if (e->n_bullets < ORC_MAX_BULLETS) {
        orc_bullet *b = &e->bullets[e->n_bullets++];
        b->x = x; b->y = y; b->dir = dir; b->owner = owner;
    }
    if (e->n_bullets > 16) e->bullet_overflow = 1;
}

/* core.py:298-326 */
int orc_spawn_bullet(orc_env *e, int pid, int dir)
{
    orc_player *pl = &e->p[pid];
    if (!pl->alive) return 0;
    if (pl->ammo <= 0) return 0;
    pl->ammo -= 1;
    append_bullet(e, pl->x, pl->y, dir, pid);
    return 1;
}

/* core.py:328-381 */
int orc_spawn_wide_shot(orc_env *e, int pid, int dir)
{
    orc_player *pl = &e->p[pid];
    if (!pl->alive) return 0;
    if (pl->ammo < WIDE_SHOT_AMMO_COST) return 0;
    pl->ammo -= WIDE_SHOT_AMMO_COST;
    int px = pl->x, py = pl->y;
    int lx[3], ly[3];
    if (dir == ORC_UP || dir == ORC_DOWN) { /* core.py:357-363: centre, x-1, x+1 */
        lx[0] = px; ly[0] = py; lx[1] = px - 1; ly[1] = py; lx[2] = px + 1; ly[2] = py;
    } else {                                /* core.py:364-370: centre, y-1, y+1 */
        lx[0] = px; ly[0] = py; lx[1] = px; ly[1] = py - 1; lx[2] = px; ly[2] = py + 1;
    }
    int spawned = 0;
    for (int i = 0; i < 3; ++i)
        if (in_bounds(e, lx[i], ly[i])) { append_bullet(e, lx[i], ly[i], dir, pid); ++spawned; }
    return spawned > 0;
}

/* core.py:383-397 */
void orc_reload_ammo(orc_env *e)
{
    for (int i = 0; i < 2; ++i) {
        orc_player *pl = &e->p[i];
        if (!pl->alive) continue;
        if (pl->ammo < MAX_AMMO) {
            pl->reload += 1;
            if (pl->reload >= RELOAD_TICKS_PER_AMMO) { pl->ammo += 1; pl->reload = 0; }
        }
    }
}

/* core.py:399-475 */
void orc_update_bullets(orc_env *e)
{
    orc_bullet moved[ORC_MAX_BULLETS];
    int n_moved = 0;
    /* phase 1 (core.py:412-435): advance, drop the ones that leave the board */
    for (int i = 0; i < e->n_bullets; ++i) {
        orc_bullet b = e->bullets[i];
        int nx = b.x + DIR_DX[b.dir], ny = b.y + DIR_DY[b.dir];
        if (!in_bounds(e, nx, ny)) continue;
        b.x = nx; b.y = ny;
        moved[n_moved++] = b;
    }
    /* phase 2 (core.py:440-473): tiles in order of first arrival (dict insertion order) */
    uint8_t grouped[ORC_MAX_BULLETS];
    memset(grouped, 0, sizeof(grouped));
    orc_bullet kept[ORC_MAX_BULLETS];
    int n_kept = 0;
    for (int i = 0; i < n_moved; ++i) {
        if (grouped[i]) continue;
        int x = moved[i].x, y = moved[i].y;
        int owner_mask = 0;
        for (int j = i; j < n_moved; ++j)
            if (!grouped[j] && moved[j].x == x && moved[j].y == y) {
                grouped[j] = 1;
                owner_mask |= 1 << moved[j].owner;
            }
        if (owner_mask == 3) continue; /* mixed owners: all gone, no flip, no hit (:444-449) */
        orc_bullet b = moved[i];       /* bullets_here[0] (:453) */
        int owner_color = PLAYER_COLOR[b.owner];
        if (get_tile(e, x, y) == owner_color) /* :458-461 */
            set_tile(e, x, y, owner_color == ORC_BLACK ? ORC_WHITE : ORC_BLACK);
        if (e->p[0].alive && b.owner != 0 && x == e->p[0].x && y == e->p[0].y) e->p[0].alive = 0;
        if (e->p[1].alive && b.owner != 1 && x == e->p[1].x && y == e->p[1].y) e->p[1].alive = 0;
        kept[n_kept++] = b; /* continues after a hit (:473) */
    }
    memcpy(e->bullets, kept, sizeof(orc_bullet) * (size_t)n_kept);
    e->n_bullets = n_kept;
}

int orc_is_round_over(const orc_env *e) { return !(e->p[0].alive && e->p[1].alive); }

/* core.py:483-495 */
int orc_get_winner(const orc_env *e)
{
    if (!orc_is_round_over(e)) return 0;
    if (!e->p[0].alive && e->p[1].alive) return 2;
    if (!e->p[1].alive && e->p[0].alive) return 1;
    return 0;
}

/* env_wrappers.py:20-66 + core.py:510-525: 0 NONE, 1-4 MOVE, 5-8 SHOOT, 9-12 CHARGE */
void orc_apply_action(orc_env *e, int pid, int a)
{
    if (!e->p[pid].alive) return;
    if (a >= 1 && a <= 4) orc_try_move(e, pid, a - 1);
    else if (a >= 5 && a <= 8) orc_spawn_bullet(e, pid, a - 5);
    else if (a >= 9 && a <= 12) orc_spawn_wide_shot(e, pid, a - 9);
}

/* core.py:497-531 */
void orc_step_players(orc_env *e, int a1, int a2)
{
    orc_apply_action(e, 0, a1);
    orc_apply_action(e, 1, a2);
    orc_reload_ammo(e);
    orc_update_bullets(e);
}

/* ------------------------------------------------------------------ wrapper: env_wrappers.py */

static int p2_can_step(const orc_env *e, int dir) /* env_wrappers.py:115-119 etc. */
{
    int nx = e->p[1].x + DIR_DX[dir], ny = e->p[1].y + DIR_DY[dir];
    return in_bounds(e, nx, ny) && get_tile(e, nx, ny) != PLAYER_COLOR[1];
}

/* env_wrappers.py:69-170 */
int orc_dummy_policy(orc_env *e)
{
    const orc_player *p1 = &e->p[0], *p2 = &e->p[1];
    e->draws_used = 0;
    if (!p2->alive) return 0; /* :77-78 */

    double move_prob, shoot_prob, random_move_prob;
    if (e->difficulty == 0) { move_prob = 0.001; shoot_prob = 0.0; random_move_prob = 0.0; }
    else                    { move_prob = 0.9;   shoot_prob = 0.2; random_move_prob = 0.05; }

    int x_al = (p2->x == p1->x), y_al = (p2->y == p1->y);
    int should_shoot = draw_random(e) < shoot_prob; /* :96 */
    if (should_shoot && p2->ammo > 0 && (x_al || y_al)) {
        if (x_al) return 5 + (p1->y < p2->y ? ORC_UP : ORC_DOWN);   /* :98-99 */
        return 5 + (p1->x < p2->x ? ORC_LEFT : ORC_RIGHT);          /* :100-101 */
    }

    int dirs[4] = {ORC_UP, ORC_DOWN, ORC_LEFT, ORC_RIGHT}; /* :104 */
    if (draw_random(e) < random_move_prob) {                /* :105 */
        draw_shuffle(e, dirs, 4);
        if (p2_can_step(e, dirs[0])) return 1 + dirs[0];
    }

    if (e->difficulty == 0) {                               /* :122-124 */
        if (draw_random(e) > move_prob) return 0;
    }

    int dx = p1->x - p2->x, dy = p1->y - p2->y;             /* :127-136 */
    int cand[2], nc = 0;
    if (dx != 0) cand[nc++] = dx > 0 ? ORC_RIGHT : ORC_LEFT;
    if (dy != 0) cand[nc++] = dy > 0 ? ORC_DOWN : ORC_UP;
    draw_shuffle(e, cand, nc);                              /* :138 */
    for (int i = 0; i < nc; ++i)
        if (p2_can_step(e, cand[i])) return 1 + cand[i];

    draw_shuffle(e, dirs, 4);                               /* :155 (list may already be permuted) */
    for (int i = 0; i < 4; ++i)
        if (p2_can_step(e, dirs[i])) return 1 + dirs[i];
    return 0;
}

/* env_wrappers.py:173-245 */
void orc_build_obs(const orc_env *e, int viewer, float *g, float *extra)
{
    const int W = e->width, H = e->height, HW = W * H;
    memset(g, 0, sizeof(float) * 12u * (size_t)HW);
    const orc_player *me = &e->p[viewer], *en = &e->p[1 - viewer];
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x)
            g[(get_tile(e, x, y) == ORC_BLACK ? 0 : 1) * HW + y * W + x] = 1.0f;
    if (me->alive && in_bounds(e, me->x, me->y)) g[2 * HW + me->y * W + me->x] = 1.0f;
    if (en->alive && in_bounds(e, en->x, en->y)) g[3 * HW + en->y * W + en->x] = 1.0f;
    for (int i = 0; i < e->n_bullets; ++i) {
        const orc_bullet *b = &e->bullets[i];
        if (!in_bounds(e, b->x, b->y)) continue;
        int ch = (b->owner == viewer ? 4 : 8) + b->dir;
        g[ch * HW + b->y * W + b->x] = 1.0f;
    }
    extra[0] = me->alive ? (float)((double)me->ammo / MAX_AMMO) : 0.0f;
    extra[1] = en->alive ? (float)((double)en->ammo / MAX_AMMO) : 0.0f;
    extra[2] = me->alive ? 1.0f : 0.0f;
    extra[3] = en->alive ? 1.0f : 0.0f;
}

/* env_wrappers.py:272-284 */
void orc_rl_reset(orc_env *e)
{
    orc_engine_reset(e);
    e->step_count = 0;
    e->episode_return = 0.0;
    e->prev_alive[0] = e->prev_alive[1] = 1;
}

static int count_white(const orc_env *e)
{
    int c = 0;
    for (int i = 0; i < e->width * e->height; ++i) c += e->grid[i] == ORC_WHITE;
    return c;
}

/* env_wrappers.py:286-444 */
int orc_rl_step(orc_env *e, int a1, int a2, orc_step_out *out)
{
    if (a1 < 0 || a1 > 12) return -1; /* :66 ValueError */
    if (e->mode == 0) a2 = orc_dummy_policy(e); /* :305-306 */
    else if (a2 < 0 || a2 > 12) return -1;

    int prev_p1_alive = e->prev_alive[0], prev_p2_alive = e->prev_alive[1]; /* :319-320 */
    int prev_white = count_white(e);                                        /* :328-329 */

    orc_step_players(e, a1, a2); /* :332 */
    e->step_count += 1;          /* :333 */

    const orc_player *p1 = &e->p[0], *p2 = &e->p[1];
    double reward = 0.0;
    int done = 0;
    memset(out, 0, sizeof(*out));
    out->a2 = a2;

    int tile_diff = count_white(e) - prev_white; /* :351-354 */
    if (tile_diff > 0) reward += tile_diff * 0.01;

    if (prev_p2_alive && !p2->alive) { reward += 1.0; out->landed_hit = 1; } /* :357-362 */
    if (prev_p1_alive && !p1->alive) { reward -= 0.01; out->got_hit = 1; }   /* :364-369 */
    if (p1->alive && p1->ammo == 0) reward -= 0.001;                         /* :372-373 */

    if (p1->alive && p2->alive) { /* :377-405 */
        int adx = p1->x - p2->x, ady = p1->y - p2->y;
        int dist = (adx < 0 ? -adx : adx) + (ady < 0 ? -ady : ady);
        int max_dist = e->width + e->height;
        double frac = (double)dist / (double)max_dist;
        double prox = 0.002 * (1.0 - frac);
        reward += prox;
        int aligned = (p1->x == p2->x) || (p1->y == p2->y);
        if (aligned) reward += 0.002;
        if (a1 >= 5 && a1 <= 12 && aligned && p1->ammo > 0) {
            int shot_dir = (a1 - 5) & 3;
            int aiming = 0;
            if (p1->x == p2->x) {
                if (p1->y < p2->y && shot_dir == ORC_DOWN) aiming = 1;
                if (p1->y > p2->y && shot_dir == ORC_UP) aiming = 1;
            } else if (p1->y == p2->y) {
                if (p1->x < p2->x && shot_dir == ORC_RIGHT) aiming = 1;
                if (p1->x > p2->x && shot_dir == ORC_LEFT) aiming = 1;
            }
            if (aiming) reward += 0.05;
        }
    }

    if (orc_is_round_over(e)) { /* :408-427 */
        done = 1;
        int w = orc_get_winner(e);
        if (w == 1) { reward += 10.0; out->win = 1; }
        else if (w == 2) { reward -= 0.1; out->lose = 1; }
    } else {
        reward -= 0.001;
    }

    e->prev_alive[0] = p1->alive; /* :430-431 */
    e->prev_alive[1] = p2->alive;

    if (e->step_count >= e->max_episode_steps) { /* :434-438 */
        done = 1;
        if (!orc_is_round_over(e)) reward -= 2.0;
    }

    e->episode_return += reward; /* :440 */
    out->reward = reward;
    out->done = done;
    out->episode_steps = e->step_count;
    out->episode_return = e->episode_return;
    return 0;
}

/* ------------------------------------------------------------------ vector runner */

int orc_max_threads(void)
{
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

int64_t orc_sizeof_env(void) { return (int64_t)sizeof(orc_env); }

static void emit_obs(const orc_env *e, int64_t i, float *obs1, float *extra1, float *obs2, float *extra2)
{
    const int64_t hw12 = 12 * (int64_t)e->width * e->height;
    float scratch_g[12 * ORC_MAX_DIM * ORC_MAX_DIM], scratch_e[4];
    if (obs1 || extra1)
        orc_build_obs(e, 0, obs1 ? obs1 + i * hw12 : scratch_g, extra1 ? extra1 + i * 4 : scratch_e);
    if (obs2 || extra2)
        orc_build_obs(e, 1, obs2 ? obs2 + i * hw12 : scratch_g, extra2 ? extra2 + i * 4 : scratch_e);
}

typedef struct {
    orc_env *envs;
    int64_t lo, hi;
    int is_step, auto_reset;
    const int8_t *a1, *a2;
    const uint32_t *table;
    float *obs1, *extra1, *obs2, *extra2, *reward;
    uint8_t *done, *flags;
    int32_t *episode_steps;
    double *episode_return;
} orc_job;

static void *run_job(void *arg)
{
    orc_job *j = (orc_job *)arg;
    for (int64_t i = j->lo; i < j->hi; ++i) {
        orc_env *e = &j->envs[i];
        e->table = j->table ? j->table + i * ORC_TABLE_STRIDE : 0;
        if (!j->is_step) {
            orc_rl_reset(e); /* env_wrappers.py:480 */
        } else {
            orc_step_out o;
            orc_rl_step(e, j->a1[i], j->a2 ? j->a2[i] : 0, &o);
            if (j->reward) j->reward[i] = (float)o.reward; /* env_wrappers.py:525 f64 -> f32 */
            if (j->done) j->done[i] = (uint8_t)o.done;
            if (j->flags)
                j->flags[i] = (uint8_t)(o.landed_hit | (o.got_hit << 1) | (o.win << 2) | (o.lose << 3));
            if (j->episode_steps) j->episode_steps[i] = o.episode_steps;
            if (j->episode_return) j->episode_return[i] = o.episode_return;
            if (o.done && j->auto_reset) orc_rl_reset(e); /* training.py:148-151 */
        }
        emit_obs(e, i, j->obs1, j->extra1, j->obs2, j->extra2);
        e->table = 0;
    }
    return 0;
}

/* plain pthread fan-out over contiguous env ranges (no OpenMP dependency) */
static void run_parallel(orc_job *proto, int64_t n, int nthreads)
{
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    if ((int64_t)nthreads > n) nthreads = n > 0 ? (int)n : 1;
    if (nthreads == 1) { proto->lo = 0; proto->hi = n; run_job(proto); return; }
    pthread_t tid[256];
    orc_job jobs[256];
    int started[256];
    for (int t = 0; t < nthreads; ++t) {
        jobs[t] = *proto;
        jobs[t].lo = n * t / nthreads;
        jobs[t].hi = n * (t + 1) / nthreads;
        started[t] = pthread_create(&tid[t], 0, run_job, &jobs[t]) == 0;
        if (!started[t]) run_job(&jobs[t]);
    }
    for (int t = 0; t < nthreads; ++t)
        if (started[t]) pthread_join(tid[t], 0);
}

/* env_wrappers.py:471-483 */
int orc_batch_reset(orc_env *envs, int64_t n, const uint32_t *table, float *obs1, float *extra1,
                    float *obs2, float *extra2, int nthreads)
{
    orc_job j;
    memset(&j, 0, sizeof(j));
    j.envs = envs; j.table = table;
    j.obs1 = obs1; j.extra1 = extra1; j.obs2 = obs2; j.extra2 = extra2;
    run_parallel(&j, n, nthreads);
    return 0;
}

/* env_wrappers.py:485-528 followed, when auto_reset, by training.py:140-151 */
int orc_batch_step(orc_env *envs, int64_t n, const int8_t *a1, const int8_t *a2,
                   const uint32_t *table, int auto_reset, float *obs1, float *extra1, float *obs2,
                   float *extra2, float *reward, uint8_t *done, uint8_t *flags,
                   int32_t *episode_steps, double *episode_return, int nthreads)
{
    for (int64_t i = 0; i < n; ++i) {
        if (a1[i] < 0 || a1[i] > 12) return -1; /* env_wrappers.py:66 */
        if (a2 && envs[i].mode == 1 && (a2[i] < 0 || a2[i] > 12)) return -1;
    }
    orc_job j;
    memset(&j, 0, sizeof(j));
    j.envs = envs; j.is_step = 1; j.auto_reset = auto_reset;
    j.a1 = a1; j.a2 = a2; j.table = table;
    j.obs1 = obs1; j.extra1 = extra1; j.obs2 = obs2; j.extra2 = extra2;
    j.reward = reward; j.done = done; j.flags = flags;
    j.episode_steps = episode_steps; j.episode_return = episode_return;
    run_parallel(&j, n, nthreads);
    return 0;
}
