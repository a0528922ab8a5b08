/*
 * stratego_oracle.h -- CPU restatement of the reference's per-step Stratego game logic.
 *
 * TEST INFRASTRUCTURE ONLY.  This is the parity oracle for the CUDA engine in
 * stratego_env_b200/.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it; the product path never does.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_golden.py checks every function here against
 * golden vectors produced by the unmodified upstream reference (oracle/gen_golden.py, run in
 * the build container against /root/reference); tests/test_oracle_vs_reference.py re-checks
 * it live against the imported reference whenever /root/reference is present.
 *
 * Algorithm source (all under /root/reference/stratego_env/):
 *   game/stratego_procedural_impl.py  ("impl")  state layout impl:16-60, 69-163
 *   stratego_multiagent_env.py        ("maenv") normalisation maenv:202-313, 387-396, 499-511
 * The dense state is the reference's own: int64[34][R][C], layers as in impl:69-96.
 */
#ifndef STRATEGO_ORACLE_H
#define STRATEGO_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { SO_NUM_STATE_LAYERS = 34, SO_PO_CHANNELS = 67, SO_FO_CHANNELS = 79,
       SO_PO_CHANNELS_ORIG = 32 /* impl:1148 */, SO_FO_CHANNELS_ORIG = 33 /* impl:1070 */ };

/* impl:253-259 */
int64_t so_action_size(int64_t R, int64_t C);
int64_t so_spatial_channels(int64_t R, int64_t C);

/* impl:213-249 */
void so_create_initial_state(int64_t R, int64_t C, const int64_t *obstacles, const int64_t *p1_map,
                             const int64_t *p2_map, int64_t max_turns, int64_t *state_out);

/* index codecs, impl:264-396 and impl:680-720.  Functions that raise in the reference return
 * a negative status instead. */
int64_t so_action_1d_from_positions(int64_t R, int64_t C, int64_t sr, int64_t sc, int64_t er, int64_t ec);
int so_spatial_from_positions(int64_t R, int64_t C, int64_t sr, int64_t sc, int64_t er, int64_t ec,
                              int64_t out_rcch[3]);
void so_positions_from_spatial(int64_t R, int64_t C, int64_t r, int64_t c, int64_t ch, int64_t out[4]);
int64_t so_action_1d_from_spatial(int64_t R, int64_t C, int64_t r, int64_t c, int64_t ch);
int so_positions_from_1d(int64_t R, int64_t C, int64_t action, int64_t out[4]);
int so_spatial_from_1d(int64_t R, int64_t C, int64_t action, int64_t out_rcch[3]);
void so_positions_from_player_perspective(int64_t R, int64_t C, int64_t player, const int64_t in[4],
                                          int64_t out[4]);
int64_t so_action_1d_from_player_perspective(int64_t R, int64_t C, int64_t action, int64_t player);

/* impl:400-517 / impl:522-642 : masks are int64, like the reference's */
void so_valid_moves_spatial_mask(int64_t R, int64_t C, const int64_t *state, int64_t player, int64_t *mask_out);
void so_valid_moves_1d_mask(int64_t R, int64_t C, const int64_t *state, int64_t player, int64_t *mask_out);

/* impl:646-675 */
void so_state_from_player_perspective(int64_t R, int64_t C, const int64_t *state, int64_t player,
                                      int64_t *out);

/* impl:726-831 */
int so_is_move_valid_by_position(int64_t R, int64_t C, const int64_t *state, int64_t player, int64_t sr,
                                 int64_t sc, int64_t er, int64_t ec, int allow_piece_oscillation);
int so_is_move_valid_by_1d_index(int64_t R, int64_t C, const int64_t *state, int64_t player, int64_t action,
                                 int allow_piece_oscillation);

/* impl:835-849 */
float so_get_game_ended(int64_t R, int64_t C, const int64_t *state, int64_t player);
int so_get_game_result_is_invalid(int64_t R, int64_t C, const int64_t *state);

/* impl:897-1045.  Returns 0, or -1 where the reference raises ValueError (state_out untouched). */
int so_get_next_state(int64_t R, int64_t C, const int64_t *state, int64_t player, int64_t action,
                      int allow_piece_oscillation, int64_t *state_out);

/* impl:854-891 : reward_matrix is float32 [13][13] */
float so_heuristic_reward(int64_t R, int64_t C, const int64_t *state, int64_t player, int64_t action,
                          const float *reward_matrix);

/* impl:1337-1397 / impl:1232-1303 : raw (un-normalised) float32 HWC observations */
void so_po_observation_ext(int64_t R, int64_t C, const int64_t *state, int64_t player, float *out);
void so_fo_observation_ext(int64_t R, int64_t C, const int64_t *state, int64_t player, float *out);

/* the deprecated "original" channel mode (obs_channel_mode='original'): impl:1153-1197 / impl:1075-1123, one
 * channel per state layer with raw values; highs and lows maenv:146-199 / maenv:87-143 */
void so_po_observation_orig(int64_t R, int64_t C, const int64_t *state, int64_t player, float *out);
void so_fo_observation_orig(int64_t R, int64_t C, const int64_t *state, int64_t player, float *out);
void so_po_highs_lows_orig(const int64_t piece_amounts[13], float highs[SO_PO_CHANNELS_ORIG], float lows[SO_PO_CHANNELS_ORIG]);
void so_fo_highs_lows_orig(const int64_t piece_amounts[13], float highs[SO_FO_CHANNELS_ORIG], float lows[SO_FO_CHANNELS_ORIG]);

/* maenv:261-313 / maenv:202-258 : per-channel highs and lows.  piece_amounts[t] for t = 1..12
 * (index 0 unused) */
void so_po_highs_lows_ext(const int64_t piece_amounts[13], float highs[SO_PO_CHANNELS], float lows[SO_PO_CHANNELS]);
void so_fo_highs_lows_ext(const int64_t piece_amounts[13], float highs[SO_FO_CHANNELS], float lows[SO_FO_CHANNELS]);
/* maenv:388-396 + maenv:499-508 : (x - mid) / range per channel, float32 arithmetic, in place */
void so_normalize(int64_t n_cells, int64_t channels, const float *highs, const float *lows, float *obs);

/* maenv:447-497 for one player: perspective flip + spatial mask + normalised obs.
 * obs_mode: 1 = partial, 2 = full, 3 = both.  Any output pointer may be NULL. */
void so_env_current_obs(int64_t R, int64_t C, const int64_t *state, int64_t player,
                        const int64_t piece_amounts[13], int obs_mode, int64_t *mask_out, float *po_out,
                        float *fo_out);

/* same with obs_channel_mode: original_channels != 0 renders the 32 / 33-channel observations */
void so_env_current_obs_ex(int64_t R, int64_t C, const int64_t *state, int64_t player,
                           const int64_t piece_amounts[13], int obs_mode, int original_channels, int64_t *mask_out,
                           float *po_out, float *fo_out);

/* maenv:659-699 : flat spatial action in the mover's frame -> next state; returns 0 / -1 (ValueError) */
int so_env_apply_spatial_action(int64_t R, int64_t C, const int64_t *state, int64_t player, int64_t flat_action,
                                int64_t *state_out);

/* CPU-baseline driver (bench.py --impl reference / cpu_baseline): random-valid self-play with the
 * loop structure of examples/basic_game_loop.py:34-63, each env re-set from setup tables when its
 * game ends.  Runs n_envs envs for steps_per_env steps each, spread over n_threads pthreads.
 * Returns the number of env-steps executed; *checksum_out folds every obs/mask so the work cannot
 * be optimised away. */
typedef struct {
    int64_t R, C, max_turns;
    int64_t usable_rows;
    int64_t piece_amounts[13];
    const int64_t *obstacles;  /* [R*C] */
    const uint8_t *setups;     /* [n_setups][usable_rows*C] own-frame piece maps, or NULL = shuffle */
    int64_t n_setups;
    int p2_rot180;             /* 1: P2 map rotated 180 deg (util:33-53); 0: row-mirrored (util:241-275) */
    int obs_mode;
} so_game_config;

int64_t so_selfplay(const so_game_config *cfg, int64_t n_envs, int64_t steps_per_env, uint64_t seed,
                    int n_threads, uint64_t *checksum_out, int64_t *games_out);

#ifdef __cplusplus
}
#endif
#endif
