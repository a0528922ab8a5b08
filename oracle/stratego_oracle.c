/*
 * stratego_oracle.c -- CPU restatement of the reference's per-step Stratego game logic.
 *
 * TEST INFRASTRUCTURE ONLY (see stratego_oracle.h).  Plain C99 + pthreads, no other dependencies.  Each function
 * cites the reference lines it restates ("impl" = stratego_env/game/stratego_procedural_impl.py,
 * "maenv" = stratego_env/stratego_multiagent_env.py, "util" = stratego_env/game/util.py).
 *
 * It deliberately keeps the reference's data model -- one dense int64[34][R][C] array per game,
 * copied on every step, flipped for player -1 -- so that it is a faithful stand-in for the
 * reference's CPU cost as well as for its results.
 */
#include "stratego_oracle.h"

#include <stdlib.h>
#include <string.h>

#include <pthread.h>

/* ---- piece codes (impl:145-163), layers (impl:69-96), recent-move codes (impl:112-125) ---- */
enum { SP_NONE = 0, SP_SPY = 1, SP_SCOUT = 2, SP_MINER = 3, SP_MARSHAL = 10, SP_FLAG = 11, SP_BOMB = 12, SP_UNKNOWN = 13 };
enum {
    L_P1 = 0, L_P2 = 1, L_OBST = 2, L_P1_PO = 3, L_P2_PO = 4, L_DATA = 5, L_P1_RECENT = 6, L_P2_RECENT = 7,
    L_P1_CAP0 = 8, L_P2_CAP0 = 20, L_P1_STILL = 32, L_P2_STILL = 33
};
enum { RM_NONE = 0, RM_CAME_FROM = 1, RM_ARRIVED = -1, RM_ARRIVED_NEXT_ILLEGAL = -2, RM_ARRIVED_CANT = -3 };

#define CELLS(R, C) ((R) * (C))
#define LAYER(state, l, R, C) ((state) + (int64_t)(l) * CELLS(R, C))
#define AT(layer, r, c, C) ((layer)[(r) * (C) + (c)])

/* scalar cells of the DATA layer, impl:136-142 */
#define D_TURN(s, R, C) (LAYER(s, L_DATA, R, C)[0 * (C) + 0])
#define D_OVER(s, R, C) (LAYER(s, L_DATA, R, C)[0 * (C) + 1])
#define D_WINNER(s, R, C) (LAYER(s, L_DATA, R, C)[0 * (C) + 2])
#define D_MAXTURNS(s, R, C) (LAYER(s, L_DATA, R, C)[1 * (C) + 0])
#define D_INVALID(s, R, C) (LAYER(s, L_DATA, R, C)[1 * (C) + 1])

/* Python floor division / modulo (numba int64 `//` and `%` floor like Python's) */
static int64_t fdiv(int64_t a, int64_t b) { int64_t q = a / b; return ((a % b != 0) && ((a < 0) != (b < 0))) ? q - 1 : q; }
static int64_t fmod_(int64_t a, int64_t b) { int64_t m = a % b; return (m != 0 && ((m < 0) != (b < 0))) ? m + b : m; }

static int own_layer(int64_t player) { return player == 1 ? L_P1 : L_P2; }           /* impl:173-176 */
static int po_layer(int64_t player) { return player == 1 ? L_P1_PO : L_P2_PO; }       /* impl:180-184 */
static int recent_layer(int64_t player) { return player == 1 ? L_P1_RECENT : L_P2_RECENT; } /* impl:188-192 */
static int still_layer(int64_t player) { return player == 1 ? L_P1_STILL : L_P2_STILL; }    /* impl:196-200 */
static int cap_layer(int64_t player, int64_t type) { return (int)((player == 1 ? 7 : 19) + type); } /* impl:99-106, 204-208 */
static int64_t mpa(int64_t R, int64_t C) { return R + C; }                            /* impl:167-169 */

int64_t so_action_size(int64_t R, int64_t C) { return R * C * mpa(R, C) + 1; }        /* impl:253-254 */
int64_t so_spatial_channels(int64_t R, int64_t C) { return (R - 1) * 2 + (C - 1) * 2 + 1; } /* impl:258-259 */

/* impl:213-249 */
void so_create_initial_state(int64_t R, int64_t C, const int64_t *obstacles, const int64_t *p1_map,
                             const int64_t *p2_map, int64_t max_turns, int64_t *st)
{
    const int64_t n = CELLS(R, C);
    memset(st, 0, sizeof(int64_t) * SO_NUM_STATE_LAYERS * n);
    for (int64_t i = 0; i < n; ++i) {
        const int64_t a = p1_map[i];
        const int64_t b = p2_map[n - 1 - i]; /* [::-1, ::-1] == reversed flat order */
        LAYER(st, L_P1, R, C)[i] = a;
        LAYER(st, L_P2, R, C)[i] = b;
        LAYER(st, L_P1_PO, R, C)[i] = a != SP_NONE ? SP_UNKNOWN : a;
        LAYER(st, L_P2_PO, R, C)[i] = b != SP_NONE ? SP_UNKNOWN : b;
        LAYER(st, L_P1_STILL, R, C)[i] = a != SP_NONE ? 1 : a;
        LAYER(st, L_P2_STILL, R, C)[i] = b != SP_NONE ? 1 : b;
        LAYER(st, L_OBST, R, C)[i] = obstacles[i];
    }
    D_MAXTURNS(st, R, C) = max_turns;
}

/* impl:264-277 */
int64_t so_action_1d_from_positions(int64_t R, int64_t C, int64_t sr, int64_t sc, int64_t er, int64_t ec)
{
    const int64_t off = (er != sr) ? er : R + ec;
    return (sr * C + sc) * mpa(R, C) + off;
}

/* impl:282-311; -1 = diagonal (assert), -2 = zero-length (ValueError) */
int so_spatial_from_positions(int64_t R, int64_t C, int64_t sr, int64_t sc, int64_t er, int64_t ec, int64_t out[3])
{
    const int64_t dc = ec - sc, dr = er - sr;
    if (!(dc == 0 || dr == 0)) return -1;
    int64_t base;
    if (dr > 0) base = 0;
    else if (dr < 0) base = R - 1;
    else if (dc > 0) base = 2 * (R - 1);
    else if (dc < 0) base = 2 * (R - 1) + (C - 1);
    else return -2;
    int64_t dist = dr + dc;
    if (dist < 0) dist = -dist;
    out[0] = sr; out[1] = sc; out[2] = base + dist - 1;
    return 0;
}

/* impl:316-335 */
void so_positions_from_spatial(int64_t R, int64_t C, int64_t r, int64_t c, int64_t ch, int64_t out[4])
{
    const int64_t mr = R - 1, mc = C - 1;
    int64_t er = r, ec = c;
    if (ch < mr) er = r + (ch + 1);
    else if (ch < 2 * mr) er = r - ((ch - mr) + 1);
    else if (ch < 2 * mr + mc) ec = c + ((ch - 2 * mr) + 1);
    else ec = c - ((ch - (2 * mr + mc)) + 1);
    out[0] = r; out[1] = c; out[2] = er; out[3] = ec;
}

/* impl:340-347 */
int64_t so_action_1d_from_spatial(int64_t R, int64_t C, int64_t r, int64_t c, int64_t ch)
{
    int64_t p[4];
    so_positions_from_spatial(R, C, r, c, ch, p);
    return so_action_1d_from_positions(R, C, p[0], p[1], p[2], p[3]);
}

/* impl:352-383; -1 = noop (ValueError) */
int so_positions_from_1d(int64_t R, int64_t C, int64_t action, int64_t out[4])
{
    if (action == so_action_size(R, C) - 1) return -1;
    const int64_t m = mpa(R, C);
    const int64_t start = fdiv(action, m);
    const int64_t sr = fdiv(start, C), sc = fmod_(start, C);
    const int64_t off = fmod_(action, m);
    int64_t er, ec;
    if (off >= R) { ec = off - R; er = sr; }
    else { er = off; ec = sc; }
    out[0] = sr; out[1] = sc; out[2] = er; out[3] = ec;
    return 0;
}

/* impl:388-396 */
int so_spatial_from_1d(int64_t R, int64_t C, int64_t action, int64_t out[3])
{
    int64_t p[4];
    if (so_positions_from_1d(R, C, action, p) != 0) return -3;
    return so_spatial_from_positions(R, C, p[0], p[1], p[2], p[3], out);
}

/* impl:680-695 */
void so_positions_from_player_perspective(int64_t R, int64_t C, int64_t player, const int64_t in[4], int64_t out[4])
{
    if (player == 1) { memcpy(out, in, 4 * sizeof(int64_t)); return; }
    out[0] = (R - 1) - in[0]; out[1] = (C - 1) - in[1];
    out[2] = (R - 1) - in[2]; out[3] = (C - 1) - in[3];
}

/* impl:700-720 */
int64_t so_action_1d_from_player_perspective(int64_t R, int64_t C, int64_t action, int64_t player)
{
    if (player == 1) return action;
    if (action == so_action_size(R, C) - 1) return action;
    int64_t p[4], f[4];
    so_positions_from_1d(R, C, action, p);
    so_positions_from_player_perspective(R, C, player, p, f);
    return so_action_1d_from_positions(R, C, f[0], f[1], f[2], f[3]);
}

/* ---- move enumeration shared by both mask forms (impl:400-517 and impl:522-642 run the same
 * scan; only the index written differs) ---- */
typedef void (*mark_fn)(int64_t R, int64_t C, int64_t sr, int64_t sc, int64_t er, int64_t ec, int64_t *mask);

static void mark_spatial(int64_t R, int64_t C, int64_t sr, int64_t sc, int64_t er, int64_t ec, int64_t *mask)
{
    int64_t i[3];
    so_spatial_from_positions(R, C, sr, sc, er, ec, i);
    mask[(i[0] * C + i[1]) * so_spatial_channels(R, C) + i[2]] = 1;
}

static void mark_1d(int64_t R, int64_t C, int64_t sr, int64_t sc, int64_t er, int64_t ec, int64_t *mask)
{
    mask[so_action_1d_from_positions(R, C, sr, sc, er, ec)] = 1;
}

/* returns 1 when at least one move was marked */
static int scan_moves(int64_t R, int64_t C, const int64_t *st, int64_t player, mark_fn mark, int64_t *mask)
{
    const int64_t *own = LAYER(st, own_layer(player), R, C);
    const int64_t *enemy = LAYER(st, own_layer(-player), R, C);
    const int64_t *obst = LAYER(st, L_OBST, R, C);
    const int64_t *recent = LAYER(st, recent_layer(player), R, C);
    static const int64_t DR[4] = {1, -1, 0, 0}, DC[4] = {0, 0, 1, -1};
    int any = 0;

    if (D_OVER(st, R, C)) return 0; /* impl:414 */

    for (int64_t sr = 0; sr < R; ++sr) {
        for (int64_t sc = 0; sc < C; ++sc) {
            const int64_t type = AT(own, sr, sc, C);
            if (type == SP_NONE || type == SP_FLAG || type == SP_BOMB) continue; /* impl:420 */
            const int64_t max_dist = (type == SP_SCOUT) ? (R > C ? R : C) : 1;    /* impl:422 / impl:492 */
            for (int d = 0; d < 4; ++d) {
                for (int64_t k = 1; k <= max_dist; ++k) {
                    const int64_t er = sr + DR[d] * k, ec = sc + DC[d] * k;
                    if (er >= R || er < 0 || ec >= C || ec < 0 || AT(obst, er, ec, C) != 0 || AT(own, er, ec, C) != 0)
                        break; /* edge, obstacle or own piece: impl:434-437, 496-499 */
                    if (AT(recent, sr, sc, C) == RM_ARRIVED_CANT && AT(recent, er, ec, C) == RM_CAME_FROM &&
                        AT(enemy, er, ec, C) == 0)
                        continue; /* two-square rule: skip this target, keep scanning: impl:439-445, 501-505 */
                    mark(R, C, sr, sc, er, ec, mask);
                    any = 1;
                    if (AT(enemy, er, ec, C) != 0) break; /* attack ends the ray: impl:454-456 */
                }
            }
        }
    }
    return any;
}

/* impl:400-517 */
void so_valid_moves_spatial_mask(int64_t R, int64_t C, const int64_t *st, int64_t player, int64_t *mask)
{
    const int64_t A = so_spatial_channels(R, C);
    memset(mask, 0, sizeof(int64_t) * R * C * A);
    if (!scan_moves(R, C, st, player, mark_spatial, mask)) mask[A - 1] = 1; /* [0,0,-1], impl:514-515 */
}

/* impl:522-642 */
void so_valid_moves_1d_mask(int64_t R, int64_t C, const int64_t *st, int64_t player, int64_t *mask)
{
    const int64_t n = so_action_size(R, C);
    memset(mask, 0, sizeof(int64_t) * n);
    if (!scan_moves(R, C, st, player, mark_1d, mask)) mask[n - 1] = 1; /* impl:639-640 */
}

static void copy_layer_rot180(int64_t *dst, const int64_t *src, int64_t n)
{
    for (int64_t i = 0; i < n; ++i) dst[i] = src[n - 1 - i];
}

/* impl:646-675 */
void so_state_from_player_perspective(int64_t R, int64_t C, const int64_t *st, int64_t player, int64_t *out)
{
    const int64_t n = CELLS(R, C);
    if (out != st) memcpy(out, st, sizeof(int64_t) * SO_NUM_STATE_LAYERS * n);
    if (player == 1) return;
    static const int pairs[][2] = {{L_P1, L_P2}, {L_P1_PO, L_P2_PO}, {L_P1_RECENT, L_P2_RECENT}, {L_P1_STILL, L_P2_STILL}};
    /* `out` may alias nothing here: the caller passes distinct buffers for player -1 */
    for (unsigned p = 0; p < sizeof(pairs) / sizeof(pairs[0]); ++p) {
        copy_layer_rot180(LAYER(out, pairs[p][0], R, C), LAYER(st, pairs[p][1], R, C), n);
        copy_layer_rot180(LAYER(out, pairs[p][1], R, C), LAYER(st, pairs[p][0], R, C), n);
    }
    copy_layer_rot180(LAYER(out, L_OBST, R, C), LAYER(st, L_OBST, R, C), n);
    for (int t = 0; t < 12; ++t) {
        copy_layer_rot180(LAYER(out, L_P1_CAP0 + t, R, C), LAYER(st, L_P2_CAP0 + t, R, C), n);
        copy_layer_rot180(LAYER(out, L_P2_CAP0 + t, R, C), LAYER(st, L_P1_CAP0 + t, R, C), n);
    }
}

static int64_t sign64(int64_t v) { return (v > 0) - (v < 0); }

/* impl:726-798 */
int so_is_move_valid_by_position(int64_t R, int64_t C, const int64_t *st, int64_t player, int64_t sr, int64_t sc,
                                 int64_t er, int64_t ec, int allow_osc)
{
    const int64_t *own = LAYER(st, own_layer(player), R, C);
    const int64_t *enemy = LAYER(st, own_layer(-player), R, C);
    const int64_t *obst = LAYER(st, L_OBST, R, C);
    const int64_t *recent = LAYER(st, recent_layer(player), R, C);

    if (D_OVER(st, R, C)) return 0;
    if (sc < 0 || sc >= C || sr < 0 || sr >= R || AT(obst, sr, sc, C) != 0) return 0;
    if (ec < 0 || ec >= C || er < 0 || er >= R || AT(obst, er, ec, C) != 0) return 0;
    const int64_t type = AT(own, sr, sc, C);
    if (type == SP_NONE || type == SP_FLAG || type == SP_BOMB) return 0;
    if (AT(own, er, ec, C) != 0) return 0;
    if (er != sr && ec != sc) return 0;
    if (AT(recent, sr, sc, C) == RM_ARRIVED_CANT && AT(recent, er, ec, C) == RM_CAME_FROM &&
        AT(enemy, er, ec, C) == 0 && !allow_osc)
        return 0;
    if (type == SP_SCOUT) {
        if (er != sr) {
            const int64_t dir = sign64(er - sr);
            for (int64_t r = sr + dir; r != er; r += dir)
                if (AT(own, r, ec, C) != 0 || AT(enemy, r, ec, C) != 0 || AT(obst, r, ec, C)) return 0;
        } else {
            const int64_t dir = sign64(ec - sc);
            /* range(sc + 0, ec, 0) cannot occur: sr == er and sc == ec was rejected above (own piece at end) */
            for (int64_t c = sc + dir; c != ec; c += dir)
                if (AT(own, er, c, C) != 0 || AT(enemy, er, c, C) != 0 || AT(obst, er, c, C)) return 0;
        }
    } else {
        int64_t a = er - sr, b = ec - sc;
        if (a < 0) a = -a;
        if (b < 0) b = -b;
        if (a > 1 || b > 1) return 0;
    }
    return 1;
}

/* impl:803-831 */
int so_is_move_valid_by_1d_index(int64_t R, int64_t C, const int64_t *st, int64_t player, int64_t action, int allow_osc)
{
    const int64_t n = so_action_size(R, C);
    if (action == n - 1) {
        int64_t *mask = (int64_t *)malloc(sizeof(int64_t) * n);
        so_valid_moves_1d_mask(R, C, st, player, mask);
        const int ok = mask[n - 1] == 1;
        free(mask);
        return ok;
    }
    int64_t p[4];
    so_positions_from_1d(R, C, action, p);
    return so_is_move_valid_by_position(R, C, st, player, p[0], p[1], p[2], p[3], allow_osc);
}

/* impl:835-842 */
float so_get_game_ended(int64_t R, int64_t C, const int64_t *st, int64_t player)
{
    if (D_OVER(st, R, C)) {
        const int64_t w = D_WINNER(st, R, C);
        if (w == 0) return (float)1e-4;
        return (float)(w * player);
    }
    return 0.0f;
}

/* impl:846-849 */
int so_get_game_result_is_invalid(int64_t R, int64_t C, const int64_t *st)
{
    return D_OVER(st, R, C) ? (D_INVALID(st, R, C) != 0) : 0;
}

/* impl:897-1045 */
int so_get_next_state(int64_t R, int64_t C, const int64_t *st, int64_t player, int64_t action, int allow_osc,
                      int64_t *ns)
{
    const int64_t n = CELLS(R, C);
    const int64_t asize = so_action_size(R, C);
    if (!so_is_move_valid_by_1d_index(R, C, st, player, action, allow_osc)) return -1; /* impl:899-902 */

    memcpy(ns, st, sizeof(int64_t) * SO_NUM_STATE_LAYERS * n); /* impl:905 */
    if (D_OVER(ns, R, C)) return 0;                            /* impl:907-909 */
    D_TURN(ns, R, C) += 1;                                     /* impl:912 */
    if (action == asize - 1) {                                 /* impl:916-920 */
        D_OVER(ns, R, C) = 1;
        D_WINNER(ns, R, C) = -player;
        return 0;
    }

    int64_t p[4];
    so_positions_from_1d(R, C, action, p);
    const int64_t sr = p[0], sc = p[1], er = p[2], ec = p[3];
    int64_t *own = LAYER(ns, own_layer(player), R, C), *enemy = LAYER(ns, own_layer(-player), R, C);
    int64_t *own_po = LAYER(ns, po_layer(player), R, C), *enemy_po = LAYER(ns, po_layer(-player), R, C);
    int64_t *own_still = LAYER(ns, still_layer(player), R, C), *enemy_still = LAYER(ns, still_layer(-player), R, C);

    AT(own_still, sr, sc, C) = 0; /* impl:939-941 */
    AT(own_still, er, ec, C) = 0;
    AT(enemy_still, er, ec, C) = 0;

    const int64_t mover = AT(own, sr, sc, C), mover_po = AT(own_po, sr, sc, C);
    const int64_t defender = AT(enemy, er, ec, C);
    AT(own, sr, sc, C) = SP_NONE; /* impl:950-951 */
    AT(own_po, sr, sc, C) = SP_NONE;

    int wins = 0, tie = 0;
    if (defender == SP_NONE) { /* impl:955-964 */
        AT(own, er, ec, C) = mover;
        int64_t a = er - sr, b = ec - sc;
        if (a < 0) a = -a;
        if (b < 0) b = -b;
        AT(own_po, er, ec, C) = (a > 1 || b > 1) ? SP_SCOUT : mover_po;
    } else { /* impl:966-995 */
        if (mover == SP_MINER && defender == SP_BOMB) wins = 1;
        else if (mover == SP_SPY && defender == SP_MARSHAL) wins = 1;
        else if (defender == SP_FLAG) { D_OVER(ns, R, C) = 1; D_WINNER(ns, R, C) = player; wins = 1; }
        else if (defender != SP_BOMB) {
            if (mover == defender) tie = 1;
            else if (mover > defender) wins = 1;
        }
        if (tie || wins) { AT(enemy, er, ec, C) = SP_NONE; AT(enemy_po, er, ec, C) = SP_NONE; }
        if (wins) { AT(own, er, ec, C) = mover; AT(own_po, er, ec, C) = mover; }
        if (!wins && !tie) AT(enemy_po, er, ec, C) = defender;
    }

    if (defender != SP_NONE) { /* impl:999-1009 */
        if (!wins) AT(LAYER(ns, cap_layer(player, mover), R, C), er, ec, C) += 1;
        if (wins || tie) AT(LAYER(ns, cap_layer(-player, defender), R, C), er, ec, C) += 1;
    }

    { /* impl:1013-1028 */
        int64_t *recent = LAYER(ns, recent_layer(player), R, C);
        const int64_t old_end = AT(recent, er, ec, C), old_start = AT(recent, sr, sc, C);
        memset(recent, 0, sizeof(int64_t) * n);
        if (defender == SP_NONE) {
            AT(recent, sr, sc, C) = RM_CAME_FROM;
            if (old_end == RM_CAME_FROM)
                AT(recent, er, ec, C) = (old_start == RM_ARRIVED_NEXT_ILLEGAL) ? RM_ARRIVED_CANT : RM_ARRIVED_NEXT_ILLEGAL;
            else
                AT(recent, er, ec, C) = RM_ARRIVED;
        }
    }

    { /* impl:1031-1036: opponent without a move loses */
        int64_t *mask = (int64_t *)malloc(sizeof(int64_t) * asize);
        so_valid_moves_1d_mask(R, C, ns, -player, mask);
        if (mask[asize - 1] == 1) { D_OVER(ns, R, C) = 1; D_WINNER(ns, R, C) = player; }
        free(mask);
    }

    if (D_TURN(ns, R, C) >= D_MAXTURNS(ns, R, C) && !D_OVER(ns, R, C)) { /* impl:1040-1043 */
        D_OVER(ns, R, C) = 1;
        D_INVALID(ns, R, C) = 1;
    }
    return 0;
}

/* impl:854-891: reward_matrix[13][13] indexed by the mover's rank at the start square and the opponent's rank at the
 * end square; the no-op scores 0; the action is assumed valid */
float so_heuristic_reward(int64_t R, int64_t C, const int64_t *state, int64_t player, int64_t action, const float *matrix)
{
    if (action == so_action_size(R, C) - 1) return 0.0f; /* impl:866-868 */
    int64_t pos[4];
    so_positions_from_1d(R, C, action, pos);             /* impl:872-874 */
    const int64_t own = AT(LAYER(state, own_layer(player), R, C), pos[0], pos[1], C);
    const int64_t enemy = AT(LAYER(state, own_layer(-player), R, C), pos[2], pos[3], C);
    return matrix[own * 13 + enemy];                     /* impl:882 */
}

/* writes plane `ch` of an HWC float tensor from an int64 layer */
static void put_plane(float *obs, int64_t n, int64_t channels, int64_t ch, const int64_t *layer)
{
    for (int64_t i = 0; i < n; ++i) obs[i * channels + ch] = (float)layer[i];
}

static void put_onehot(float *obs, int64_t n, int64_t channels, int64_t ch0, int64_t n_types, const int64_t *layer)
{
    for (int64_t t = 1; t <= n_types; ++t)
        for (int64_t i = 0; i < n; ++i) obs[i * channels + ch0 + (t - 1)] = (layer[i] == t) ? 1.0f : 0.0f;
}

/* impl:1337-1397, channel map impl:1306-1332 */
void so_po_observation_ext(int64_t R, int64_t C, const int64_t *state, int64_t player, float *obs)
{
    const int64_t n = CELLS(R, C), ch = SO_PO_CHANNELS;
    int64_t *buf = NULL;
    const int64_t *st = state; /* player 1: same array, impl:647-648 */
    if (player != 1) {
        buf = (int64_t *)malloc(sizeof(int64_t) * SO_NUM_STATE_LAYERS * n);
        so_state_from_player_perspective(R, C, state, player, buf); /* impl:1338 */
        st = buf;
    }
    put_onehot(obs, n, ch, 0, 12, LAYER(st, L_P1, R, C));
    put_onehot(obs, n, ch, 12, 13, LAYER(st, L_P1_PO, R, C));
    put_onehot(obs, n, ch, 25, 13, LAYER(st, L_P2_PO, R, C));
    put_plane(obs, n, ch, 38, LAYER(st, L_OBST, R, C));
    put_plane(obs, n, ch, 39, LAYER(st, L_P1_RECENT, R, C));
    put_plane(obs, n, ch, 40, LAYER(st, L_P2_RECENT, R, C));
    for (int t = 0; t < 12; ++t) put_plane(obs, n, ch, 41 + t, LAYER(st, L_P1_CAP0 + t, R, C));
    for (int t = 0; t < 12; ++t) put_plane(obs, n, ch, 53 + t, LAYER(st, L_P2_CAP0 + t, R, C));
    put_plane(obs, n, ch, 65, LAYER(st, L_P1_STILL, R, C));
    put_plane(obs, n, ch, 66, LAYER(st, L_P2_STILL, R, C));
    free(buf);
}

/* impl:1232-1303, channel map impl:1200-1227 */
void so_fo_observation_ext(int64_t R, int64_t C, const int64_t *state, int64_t player, float *obs)
{
    const int64_t n = CELLS(R, C), ch = SO_FO_CHANNELS;
    int64_t *buf = NULL;
    const int64_t *st = state; /* player 1: same array, impl:647-648 */
    if (player != 1) {
        buf = (int64_t *)malloc(sizeof(int64_t) * SO_NUM_STATE_LAYERS * n);
        so_state_from_player_perspective(R, C, state, player, buf); /* impl:1233 */
        st = buf;
    }
    put_onehot(obs, n, ch, 0, 12, LAYER(st, L_P1, R, C));
    put_onehot(obs, n, ch, 12, 12, LAYER(st, L_P2, R, C));
    put_onehot(obs, n, ch, 24, 13, LAYER(st, L_P1_PO, R, C));
    put_onehot(obs, n, ch, 37, 13, LAYER(st, L_P2_PO, R, C));
    put_plane(obs, n, ch, 50, LAYER(st, L_OBST, R, C));
    put_plane(obs, n, ch, 51, LAYER(st, L_P1_RECENT, R, C));
    put_plane(obs, n, ch, 52, LAYER(st, L_P2_RECENT, R, C));
    for (int t = 0; t < 12; ++t) put_plane(obs, n, ch, 53 + t, LAYER(st, L_P1_CAP0 + t, R, C));
    for (int t = 0; t < 12; ++t) put_plane(obs, n, ch, 65 + t, LAYER(st, L_P2_CAP0 + t, R, C));
    put_plane(obs, n, ch, 77, LAYER(st, L_P1_STILL, R, C));
    put_plane(obs, n, ch, 78, LAYER(st, L_P2_STILL, R, C));
    free(buf);
}

/* Deprecated "original" channel mode (obs_channel_mode='original', maenv:370-375): raw layer values, one
 * channel per state layer.  impl:1153-1197, channel map impl:1126-1148 */
void so_po_observation_orig(int64_t R, int64_t C, const int64_t *state, int64_t player, float *obs)
{
    const int64_t n = CELLS(R, C), ch = SO_PO_CHANNELS_ORIG;
    int64_t *buf = NULL;
    const int64_t *st = state;
    if (player != 1) {
        buf = (int64_t *)malloc(sizeof(int64_t) * SO_NUM_STATE_LAYERS * n);
        so_state_from_player_perspective(R, C, state, player, buf); /* impl:1156 */
        st = buf;
    }
    put_plane(obs, n, ch, 0, LAYER(st, L_P1, R, C));
    put_plane(obs, n, ch, 1, LAYER(st, L_P1_PO, R, C));
    put_plane(obs, n, ch, 2, LAYER(st, L_P2_PO, R, C));
    put_plane(obs, n, ch, 3, LAYER(st, L_OBST, R, C));
    put_plane(obs, n, ch, 4, LAYER(st, L_P1_RECENT, R, C));
    put_plane(obs, n, ch, 5, LAYER(st, L_P2_RECENT, R, C));
    for (int t = 0; t < 12; ++t) put_plane(obs, n, ch, 6 + t, LAYER(st, L_P1_CAP0 + t, R, C));
    for (int t = 0; t < 12; ++t) put_plane(obs, n, ch, 18 + t, LAYER(st, L_P2_CAP0 + t, R, C));
    put_plane(obs, n, ch, 30, LAYER(st, L_P1_STILL, R, C));
    put_plane(obs, n, ch, 31, LAYER(st, L_P2_STILL, R, C));
    free(buf);
}

/* impl:1075-1123, channel map impl:1048-1070 */
void so_fo_observation_orig(int64_t R, int64_t C, const int64_t *state, int64_t player, float *obs)
{
    const int64_t n = CELLS(R, C), ch = SO_FO_CHANNELS_ORIG;
    int64_t *buf = NULL;
    const int64_t *st = state;
    if (player != 1) {
        buf = (int64_t *)malloc(sizeof(int64_t) * SO_NUM_STATE_LAYERS * n);
        so_state_from_player_perspective(R, C, state, player, buf); /* impl:1078 */
        st = buf;
    }
    put_plane(obs, n, ch, 0, LAYER(st, L_P1, R, C));
    put_plane(obs, n, ch, 1, LAYER(st, L_P2, R, C));
    put_plane(obs, n, ch, 2, LAYER(st, L_OBST, R, C));
    put_plane(obs, n, ch, 3, LAYER(st, L_P1_RECENT, R, C));
    put_plane(obs, n, ch, 4, LAYER(st, L_P2_RECENT, R, C));
    put_plane(obs, n, ch, 5, LAYER(st, L_P1_PO, R, C));
    put_plane(obs, n, ch, 6, LAYER(st, L_P2_PO, R, C));
    for (int t = 0; t < 12; ++t) put_plane(obs, n, ch, 7 + t, LAYER(st, L_P1_CAP0 + t, R, C));
    for (int t = 0; t < 12; ++t) put_plane(obs, n, ch, 19 + t, LAYER(st, L_P2_CAP0 + t, R, C));
    put_plane(obs, n, ch, 31, LAYER(st, L_P1_STILL, R, C));
    put_plane(obs, n, ch, 32, LAYER(st, L_P2_STILL, R, C));
    free(buf);
}

/* maenv:146-199 */
void so_po_highs_lows_orig(const int64_t amounts[13], float hi[SO_PO_CHANNELS_ORIG], float lo[SO_PO_CHANNELS_ORIG])
{
    for (int c = 0; c < SO_PO_CHANNELS_ORIG; ++c) { hi[c] = 2.0f; lo[c] = 0.0f; } /* obstacles, captured, still */
    hi[0] = (float)SP_BOMB;                                                       /* own true ranks */
    hi[1] = hi[2] = (float)SP_UNKNOWN;                                            /* PO ranks */
    hi[4] = hi[5] = (float)RM_CAME_FROM; lo[4] = lo[5] = (float)RM_ARRIVED_CANT;
    for (int t = 1; t <= 12; ++t)
        if (amounts[t] > 1) { hi[6 + t - 1] = (float)amounts[t]; hi[18 + t - 1] = (float)amounts[t]; } /* maenv:176-180 */
}

/* maenv:87-143 */
void so_fo_highs_lows_orig(const int64_t amounts[13], float hi[SO_FO_CHANNELS_ORIG], float lo[SO_FO_CHANNELS_ORIG])
{
    for (int c = 0; c < SO_FO_CHANNELS_ORIG; ++c) { hi[c] = 2.0f; lo[c] = 0.0f; }
    hi[0] = hi[1] = (float)SP_BOMB;
    hi[3] = hi[4] = (float)RM_CAME_FROM; lo[3] = lo[4] = (float)RM_ARRIVED_CANT;
    hi[5] = hi[6] = (float)SP_UNKNOWN;
    for (int t = 1; t <= 12; ++t)
        if (amounts[t] > 1) { hi[7 + t - 1] = (float)amounts[t]; hi[19 + t - 1] = (float)amounts[t]; } /* maenv:120-124 */
}

/* maenv:261-313 */
void so_po_highs_lows_ext(const int64_t amounts[13], float hi[SO_PO_CHANNELS], float lo[SO_PO_CHANNELS])
{
    for (int c = 0; c < 38; ++c) { hi[c] = 1.0f; lo[c] = -1.0f; }  /* one-hot groups 0-37 */
    hi[38] = 1.0f; lo[38] = -1.0f;                                  /* obstacles */
    hi[39] = hi[40] = (float)RM_CAME_FROM; lo[39] = lo[40] = (float)RM_ARRIVED_CANT;
    for (int c = 41; c < 65; ++c) { hi[c] = 8.0f; lo[c] = 0.0f; }   /* captured counts */
    hi[65] = hi[66] = 1.0f; lo[65] = lo[66] = -1.0f;                /* still flags */
    for (int t = 1; t <= 12; ++t)
        if (amounts[t] > 1) { hi[41 + t - 1] = (float)amounts[t]; hi[53 + t - 1] = (float)amounts[t]; } /* maenv:294-298 */
}

/* maenv:202-258 */
void so_fo_highs_lows_ext(const int64_t amounts[13], float hi[SO_FO_CHANNELS], float lo[SO_FO_CHANNELS])
{
    for (int c = 0; c < 51; ++c) { hi[c] = 1.0f; lo[c] = -1.0f; }
    hi[51] = hi[52] = (float)RM_CAME_FROM; lo[51] = lo[52] = (float)RM_ARRIVED_CANT;
    for (int c = 53; c < 77; ++c) { hi[c] = 8.0f; lo[c] = 0.0f; }
    hi[77] = hi[78] = 1.0f; lo[77] = lo[78] = -1.0f;
    for (int t = 1; t <= 12; ++t)
        if (amounts[t] > 1) { hi[53 + t - 1] = (float)amounts[t]; hi[65 + t - 1] = (float)amounts[t]; } /* maenv:238-242 */
}

/* maenv:388-396 (ranges, mids) and maenv:499-508 ((x - mid) / range); all float32 */
void so_normalize(int64_t n_cells, int64_t channels, const float *hi, const float *lo, float *obs)
{
    for (int64_t i = 0; i < n_cells; ++i) {
        for (int64_t c = 0; c < channels; ++c) {
            const float range = (hi[c] - lo[c]) / 2.0f; /* built with -ffp-contract=off, no fast-math */
            const float mid = (hi[c] + lo[c]) / 2.0f;
            const float d = obs[i * channels + c] - mid;
            obs[i * channels + c] = d / range;
        }
    }
}

/* maenv:447-497; original_channels selects obs_channel_mode='original' (maenv:370-375, 460-468, 479-487) */
void so_env_current_obs_ex(int64_t R, int64_t C, const int64_t *state, int64_t player, const int64_t amounts[13],
                           int obs_mode, int original_channels, int64_t *mask_out, float *po_out, float *fo_out)
{
    const int64_t n = CELLS(R, C);
    int64_t *persp = (int64_t *)malloc(sizeof(int64_t) * SO_NUM_STATE_LAYERS * n);
    so_state_from_player_perspective(R, C, state, player, persp);       /* maenv:452 */
    if (mask_out) so_valid_moves_spatial_mask(R, C, persp, 1, mask_out); /* maenv:454 */
    if ((obs_mode & 1) && po_out) {                                     /* maenv:458-475 */
        float hi[SO_PO_CHANNELS], lo[SO_PO_CHANNELS];
        if (original_channels) {
            so_po_observation_orig(R, C, persp, 1, po_out);
            so_po_highs_lows_orig(amounts, hi, lo);
            so_normalize(n, SO_PO_CHANNELS_ORIG, hi, lo, po_out);
        } else {
            so_po_observation_ext(R, C, persp, 1, po_out);
            so_po_highs_lows_ext(amounts, hi, lo);
            so_normalize(n, SO_PO_CHANNELS, hi, lo, po_out);
        }
    }
    if ((obs_mode & 2) && fo_out) {                                     /* maenv:477-492 */
        float hi[SO_FO_CHANNELS], lo[SO_FO_CHANNELS];
        if (original_channels) {
            so_fo_observation_orig(R, C, persp, 1, fo_out);
            so_fo_highs_lows_orig(amounts, hi, lo);
            so_normalize(n, SO_FO_CHANNELS_ORIG, hi, lo, fo_out);
        } else {
            so_fo_observation_ext(R, C, persp, 1, fo_out);
            so_fo_highs_lows_ext(amounts, hi, lo);
            so_normalize(n, SO_FO_CHANNELS, hi, lo, fo_out);
        }
    }
    free(persp);
}

void so_env_current_obs(int64_t R, int64_t C, const int64_t *state, int64_t player, const int64_t amounts[13],
                        int obs_mode, int64_t *mask_out, float *po_out, float *fo_out)
{
    so_env_current_obs_ex(R, C, state, player, amounts, obs_mode, 0, mask_out, po_out, fo_out);
}

/* maenv:684-692 */
int so_env_apply_spatial_action(int64_t R, int64_t C, const int64_t *state, int64_t player, int64_t flat, int64_t *ns)
{
    const int64_t A = so_spatial_channels(R, C);
    /* np.unravel_index(flat, (R, C, A)), maenv:685 */
    const int64_t ch = flat % A, cell = flat / A;
    const int64_t r = cell / C, c = cell % C;
    if (flat < 0 || r >= R) return -1;
    int64_t a = so_action_1d_from_spatial(R, C, r, c, ch);             /* maenv:686 */
    a = so_action_1d_from_player_perspective(R, C, a, player);         /* maenv:689 */
    return so_get_next_state(R, C, state, player, a, 0, ns);           /* maenv:691 */
}

/* ---------------------------------------------------------------------------------------------
 * CPU-baseline self-play driver.  Loop structure of examples/basic_game_loop.py:34-63 with a
 * fast uniform sampler over the valid-action mask (the as-shipped np.random.choice sampler,
 * maenv:830-834, costs 4x the env step and is not part of the path being measured).
 * ------------------------------------------------------------------------------------------- */
static uint64_t splitmix64(uint64_t *s)
{
    uint64_t z = (*s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

/* util:13-30 (shuffle of the usable cells) or util:241-275 (row of the human-setup table) */
static void draw_piece_map(const so_game_config *cfg, uint64_t *rng, int64_t *map)
{
    const int64_t n = cfg->R * cfg->C, m = cfg->usable_rows * cfg->C;
    memset(map, 0, sizeof(int64_t) * n);
    if (cfg->setups) {
        const uint8_t *row = cfg->setups + (splitmix64(rng) % (uint64_t)cfg->n_setups) * m;
        for (int64_t i = 0; i < m; ++i) map[i] = row[i];
        return;
    }
    int64_t cells[256];
    for (int64_t i = 0; i < m; ++i) cells[i] = i;
    for (int64_t i = m - 1; i >= 1; --i) {
        const int64_t j = (int64_t)(splitmix64(rng) % (uint64_t)(i + 1));
        const int64_t t = cells[i]; cells[i] = cells[j]; cells[j] = t;
    }
    int64_t k = 0;
    for (int t = 1; t <= 12; ++t)
        for (int64_t a = 0; a < cfg->piece_amounts[t]; ++a) map[cells[k++]] = t;
}

static void new_game(const so_game_config *cfg, uint64_t *rng, int64_t *state, int64_t *m1, int64_t *m2)
{
    const int64_t R = cfg->R, C = cfg->C;
    draw_piece_map(cfg, rng, m1);
    draw_piece_map(cfg, rng, m2);
    if (!cfg->p2_rot180) { /* human tables: P2 ends up row-mirrored only, so pre-mirror the columns */
        for (int64_t r = 0; r < R; ++r)
            for (int64_t c = 0; c < C / 2; ++c) {
                const int64_t t = m2[r * C + c]; m2[r * C + c] = m2[r * C + (C - 1 - c)]; m2[r * C + (C - 1 - c)] = t;
            }
    }
    so_create_initial_state(R, C, cfg->obstacles, m1, m2, cfg->max_turns, state);
}

typedef struct {
    const so_game_config *cfg;
    int64_t n_envs, steps_per_env;
    uint64_t seed;
    int64_t *next_env; /* shared work counter */
    uint64_t checksum;
    int64_t total, games;
} selfplay_job;

static void *selfplay_worker(void *arg)
{
    selfplay_job *job = (selfplay_job *)arg;
    const so_game_config *cfg = job->cfg;
    const int64_t R = cfg->R, C = cfg->C, n = R * C, A = so_spatial_channels(R, C);
    int64_t *state = (int64_t *)malloc(sizeof(int64_t) * SO_NUM_STATE_LAYERS * n);
    int64_t *next = (int64_t *)malloc(sizeof(int64_t) * SO_NUM_STATE_LAYERS * n);
    int64_t *m1 = (int64_t *)malloc(sizeof(int64_t) * n), *m2 = (int64_t *)malloc(sizeof(int64_t) * n);
    int64_t *mask = (int64_t *)malloc(sizeof(int64_t) * n * A);
    float *po = (float *)malloc(sizeof(float) * n * SO_PO_CHANNELS);
    float *fo = (float *)malloc(sizeof(float) * n * SO_FO_CHANNELS);
    for (;;) {
        const int64_t e = __atomic_fetch_add(job->next_env, 1, __ATOMIC_RELAXED);
        if (e >= job->n_envs) break;
        uint64_t rng = job->seed * 0x9E3779B97F4A7C15ull + (uint64_t)e * 0xD1B54A32D192ED03ull + 1;
        uint64_t local_sum = 0;
        int64_t player = 1;
        new_game(cfg, &rng, state, m1, m2);
        so_env_current_obs(R, C, state, player, cfg->piece_amounts, cfg->obs_mode, mask, po, fo); /* reset(), maenv:621 */
        for (int64_t s = 0; s < job->steps_per_env; ++s) {
            /* uniform draw over the valid entries of the mask */
            int64_t count = 0;
            for (int64_t i = 0; i < n * A; ++i) count += mask[i];
            int64_t k = (int64_t)(splitmix64(&rng) % (uint64_t)count), action = -1;
            for (int64_t i = 0; i < n * A; ++i)
                if (mask[i] && k-- == 0) { action = i; break; }
            if (so_env_apply_spatial_action(R, C, state, player, action, next) != 0) {
                /* only reachable for a noop-only mask on a live game; start over */
                new_game(cfg, &rng, state, m1, m2);
                player = 1;
            } else {
                int64_t *tmp = state; state = next; next = tmp;
                player = -player; /* penv:153 */
            }
            const float reward = so_get_game_ended(R, C, state, player); /* maenv:699 */
            if (reward != 0.0f) {
                /* terminal: the reference renders both players' observations, maenv:772-773 */
                so_env_current_obs(R, C, state, 1, cfg->piece_amounts, cfg->obs_mode, mask, po, fo);
                so_env_current_obs(R, C, state, -1, cfg->piece_amounts, cfg->obs_mode, mask, po, fo);
                job->games += 1;
                new_game(cfg, &rng, state, m1, m2);
                player = 1;
            }
            so_env_current_obs(R, C, state, player, cfg->piece_amounts, cfg->obs_mode, mask, po, fo); /* maenv:768 */
            if (cfg->obs_mode & 1) local_sum += (uint64_t)(po[(s * 7) % (n * SO_PO_CHANNELS)] * 4.0f + 8.0f);
            if (cfg->obs_mode & 2) local_sum += (uint64_t)(fo[(s * 7) % (n * SO_FO_CHANNELS)] * 4.0f + 8.0f);
            local_sum = local_sum * 31 + (uint64_t)action;
            job->total += 1;
        }
        job->checksum ^= local_sum;
    }
    free(state); free(next); free(m1); free(m2); free(mask); free(po); free(fo);
    return NULL;
}

int64_t so_selfplay(const so_game_config *cfg, int64_t n_envs, int64_t steps_per_env, uint64_t seed, int n_threads,
                    uint64_t *checksum_out, int64_t *games_out)
{
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 1024) n_threads = 1024;
    pthread_t *threads = (pthread_t *)malloc(sizeof(pthread_t) * n_threads);
    selfplay_job *jobs = (selfplay_job *)calloc(n_threads, sizeof(selfplay_job));
    int64_t next_env = 0;
    for (int t = 0; t < n_threads; ++t) {
        jobs[t].cfg = cfg; jobs[t].n_envs = n_envs; jobs[t].steps_per_env = steps_per_env;
        jobs[t].seed = seed; jobs[t].next_env = &next_env;
        if (t > 0) pthread_create(&threads[t], NULL, selfplay_worker, &jobs[t]);
    }
    selfplay_worker(&jobs[0]);
    uint64_t checksum = jobs[0].checksum;
    int64_t total = jobs[0].total, games = jobs[0].games;
    for (int t = 1; t < n_threads; ++t) {
        pthread_join(threads[t], NULL);
        checksum ^= jobs[t].checksum; total += jobs[t].total; games += jobs[t].games;
    }
    free(threads); free(jobs);
    if (checksum_out) *checksum_out = checksum;
    if (games_out) *games_out = games;
    return total;
}
