/*
 * stratego_b200.h -- C ABI of the B200-native batched Stratego engine.
 *
 * This is the drop-in boundary for the hot path of JBLanier/stratego_env: the per-step game logic
 * of stratego_env/game (valid-action mask, move/combat resolution, observation rendering,
 * auto-reset).  The reference has no FFI of its own -- the path sits behind two plain-Python
 * interfaces, StrategoProceduralEnv (stratego_env/game/stratego_procedural_env.py:20-173, "penv")
 * and StrategoMultiAgentEnv (stratego_env/stratego_multiagent_env.py:316-834, "maenv") -- so each
 * entry point below cites the reference function(s) it replaces ("impl" =
 * stratego_env/game/stratego_procedural_impl.py).  INTEGRATION.md shows the ctypes binding.
 *
 * Conventions
 *   - Every `*_d` / sx_state pointer is a DEVICE pointer owned by the caller (torch tensors in the
 *     Python host layer).  The library never allocates or frees caller-visible memory, except for
 *     the sx_host_env convenience object at the bottom, which owns its own device buffers.
 *   - Every call enqueues work on `stream` (a cudaStream_t passed as void*) and returns without
 *     synchronising.  Return value: 0 = ok, negative = error (sx_last_error() has the text).
 *   - Games are independent: `num_envs` games, game b uses row b of every tensor.
 *   - "player" is +1 (moves first, owns rows 0..) or -1, as in the reference.
 *   - Frames: observations, spatial masks and spatial actions are in the frame of the player they
 *     are for (player -1 sees the board rotated 180 degrees, impl:646-675); device state and 1D
 *     actions are in the absolute frame, as in the reference.
 */
#ifndef STRATEGO_B200_H
#define STRATEGO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SX_PO_CHANNELS 67 /* impl:1332 */
#define SX_FO_CHANNELS 79 /* impl:1227 */
#define SX_PO_CHANNELS_ORIGINAL 32 /* impl:1148, deprecated obs_channel_mode='original' */
#define SX_FO_CHANNELS_ORIGINAL 33 /* impl:1070 */
#define SX_NUM_STATE_LAYERS 34 /* impl:109 */
#define SX_MAX_CAPTURE_COUNT 8

typedef struct sx_config sx_config; /* opaque, host-side */

/* Game variant description (one of the reference's config.py:3-313 dicts plus the host-computed
 * normalisation tables, maenv:202-313, 387-396, 499-508).  All pointers are HOST pointers and are
 * copied. */
typedef struct {
    int32_t rows, cols;           /* 3..15 each */
    int32_t max_turns;            /* 1..65535 */
    int32_t usable_rows;          /* initial_state_usable_rows */
    int32_t piece_amounts[13];    /* per piece code 1..12 (0 unused) */
    const uint8_t *obstacles;     /* [rows*cols] 0/1 */
    const float *captured_lut;    /* [12][SX_MAX_CAPTURE_COUNT+1] normalised captured-count values */
    const float *recent_lut;      /* [5] normalised recent-move codes -3..+1 */
    const float *unit_lut;        /* [2] normalised 0 and 1 of the one-hot/obstacle/still channels */
    int32_t p2_rot180;            /* setup rows for player -1 are rotated 180 deg (util:33-53) instead of
                                     row-mirrored (human tables, util:241-275) */
    int32_t capture_capacity;     /* capture-list entries per game; 0 = 2 * pieces per side (enough for any game
                                     played from this variant's setups); the stateless facade passes rows*cols */
    int32_t obs_channel_mode;     /* SX_CHANNELS_EXTENDED (67 / 79 one-hot channels, impl:1232-1397) or
                                     SX_CHANNELS_ORIGINAL (deprecated 32 / 33 raw-value channels, impl:1048-1197;
                                     env_config['obs_channel_mode'], maenv:370-375).  In the original mode unit_lut
                                     holds the normalised 0 and 1 of the obstacle / still channels (hi 2, lo 0) and
                                     captured_lut the hi-2 variant (maenv:87-199) */
    const float *rank_lut;        /* [14] original mode: normalised true rank 0..12 (entry 13 unused) */
    const float *po_rank_lut;     /* [14] original mode: normalised partially observable rank 0..13 */
} sx_config_desc;

enum { SX_CHANNELS_EXTENDED = 0, SX_CHANNELS_ORIGINAL = 1 };

/* Byte/element strides of the device tensors for one variant. */
typedef struct {
    int32_t rows, cols, cells;
    int32_t spatial_channels;     /* A = 2(R-1)+2(C-1)+1, impl:258-259 */
    int32_t spatial_actions;      /* R*C*A = bytes of one uint8 spatial mask */
    int32_t action_size;          /* R*C*(R+C)+1, impl:253-254 */
    int32_t board_stride;         /* uint8 per env (cells rounded up to 16) */
    int32_t aux_stride;           /* int16 per env (= 8) */
    int32_t captured_stride;      /* uint16 per env */
    int32_t po_floats, fo_floats; /* float32 per env of each observation */
    int32_t setup_len;            /* usable_rows*cols bytes per setup-table row */
    int32_t pieces_per_side;
    int32_t po_channels, fo_channels; /* 67 / 79, or 32 / 33 in the original channel mode */
} sx_layout;

/* Compact struct-of-arrays device state (DESIGN.md "State layout"). */
typedef struct {
    uint8_t *board;     /* [num_envs][board_stride] one packed byte per cell */
    int16_t *aux;       /* [num_envs][8] turn, max_turns, flags, recent moves, episode counter */
    uint16_t *captured; /* [num_envs][captured_stride] (cell, owner, type, count) capture entries */
} sx_state;

/* Per-step outputs; any pointer may be NULL to skip that output. */
typedef struct {
    float *partial_obs;      /* [num_envs][R][C][67 | 32] float32, maenv:461-475 (normalised) */
    float *full_obs;         /* [num_envs][R][C][79 | 33] float32, maenv:480-492 (normalised) */
    uint8_t *valid_mask;     /* [num_envs][R][C][A] uint8 0/1, maenv:454 / impl:400-517 */
    float *reward;           /* [num_envs] player +1's reward when the game ended this step, else 0
                                (maenv:777-801: +-1, or 0 for an invalid ending) */
    uint8_t *done;           /* [num_envs] game ended this step (maenv:772) */
    int8_t *winner;          /* [num_envs] +1 / -1 / 0, state[5,0,2] (impl:140) of the ended game */
    uint8_t *ending_invalid; /* [num_envs] impl:846-849 */
    uint8_t *illegal;        /* [num_envs] 1 = action rejected, game left untouched (ValueError, impl:899-902);
                                2 = the game's capture list overflowed: the compact state no longer matches what the
                                reference's dense counters (impl:999-1009) would hold (sticky until the game is re-set;
                                cannot happen in games played from the variant's own setups) */
    int8_t *player;          /* [num_envs] player the returned mask/obs are for (= player to move) */
    int32_t *next_action;    /* [num_envs] uniformly sampled valid spatial action for `player`
                                (replaces maenv:830-834), written when SX_SAMPLE_NEXT is set */
    /* maenv:772-773: when a game ends, BOTH players get their observation of the final position.  With SX_AUTO_RESET the
     * regular outputs already show the next game, so these optional side buffers receive the terminal observations:
     * [num_envs][2][R][C][channels], index 0 = player +1's view, 1 = player -1's; only rows of games with done = 1 are
     * written.  (The terminal valid-action mask is the lone noop entry [0,0,A-1] for both players, impl:414 / 514-515.)
     * Each needs its regular counterpart (partial_obs / full_obs) to be requested in the same call. */
    float *terminal_partial_obs;
    float *terminal_full_obs;
} sx_outputs;

enum {
    SX_ACTION_SPATIAL = 0, /* flat index into (R,C,A) in the mover's frame, as fed to maenv.step (maenv:685) */
    SX_ACTION_1D = 1       /* absolute 1D index incl. trailing noop, as fed to penv.get_next_state (penv:148) */
};

enum {
    SX_AUTO_RESET = 1,          /* games that end are re-set in the same call; outputs show the new game */
    SX_SAMPLE_NEXT = 2,         /* also draw a uniformly random valid action into outputs.next_action */
    SX_ALLOW_OSCILLATION = 4,   /* allow_piece_oscillation=True (impl:771-777) */
    SX_RESET_RANDOM_SHUFFLE = 8, /* (re)sets draw setups by shuffling the pieces (util:13-30) instead of a table */
    SX_SAME_SETUP = 32,         /* same_start_pos_everytime (maenv:352-354): every game of an env starts from the setup
                                   its first draw produced */
    SX_REPEAT_OTHER_SIDE = 64,  /* repeat_games_from_other_side (maenv:530-534): every second game of an env repeats the
                                   previous game's initial position as player -1 sees it (impl:646-675), player -1 first */
    SX_KERNEL_BASELINE = 16     /* run the general warp-per-game kernel even where a specialised one is eligible (the
                                   thread-per-game kernel of the <= 16-cell boards).  Results are identical by
                                   contract; the cross-kernel parity tests use it */
};

const char *sx_last_error(void);
int sx_version(void);

int sx_config_create(const sx_config_desc *desc, sx_config **out);
void sx_config_destroy(sx_config *cfg);
int sx_config_layout(const sx_config *cfg, sx_layout *out);

/* Launch tuning of the warp-level kernel; a development aid (tools/sweep_fused.py times the shipped library with it).
 * Every setting gives bit-identical results -- only the speed changes.  -1 keeps the built-in choice.
 *   warps_per_block  resident warps per SM (0 = built-in again)
 *   issue_point      where a game's background copy is handed to the TMA engine: 0 right before the sparse entries,
 *                    1 after the outcome of the move is known, 2 at the top of the game
 *   compact_movers   0 / 1: move generation hands the movable pieces to lanes instead of walking cells (10x10 class) */
int sx_config_set_tuning(sx_config *cfg, int32_t warps_per_block, int32_t issue_point, int32_t compact_movers);

/* Curriculum start states (curriculum_start_states_path, maenv:341-351 / 519-527, util:373-387): from now on every
 * (re)set of a game with this configuration -- sx_reset and the auto-reset inside sx_step_all -- copies a uniformly
 * drawn entry of `table` (n_states compact states, e.g. made with sx_import_ref_state from the file's dense states)
 * instead of dealing setups; the turn counter restarts at 0 with the configured max_turns (util:382-383) and the player
 * to move is drawn uniformly (maenv:523).  The table must stay allocated while it is set.  n_states = 0 (or a NULL
 * board pointer) switches back to setups.
 * start_index_d (optional, device, int32 [index_len]): start_index_d[global env id - index_env_base] receives the entry
 * every game was (re)started from -- maenv:524-527 maps the players through the entry's likely winner.  Global env id =
 * env_base + the env's position in the call, as everywhere else; ids outside [index_env_base, index_env_base +
 * index_len) are not recorded. */
int sx_config_set_start_states(sx_config *cfg, sx_state table, int64_t n_states, int32_t *start_index_d,
                               int64_t index_env_base, int64_t index_len);

/* Replaces penv.create_initial_state (penv:38 -> impl:213-249) + the setup samplers
 * (util:33-53 random, util:301-319 human).  Re-sets every env b with reset_mask_d[b] != 0 (all when
 * NULL).  Setups come from `setups_d` ([n_setups][setup_len] own-frame piece maps): rows
 * setup_idx_d[b][0..1] when given, else two Philox draws keyed by (seed, env_base + b, episode).
 * With SX_RESET_RANDOM_SHUFFLE in flags the table is ignored and pieces are shuffled on device. */
int sx_reset(const sx_config *cfg, sx_state st, int64_t num_envs, int64_t env_base, const uint8_t *reset_mask_d,
             const uint8_t *setups_d, int32_t n_setups, const int32_t *setup_idx_d, uint64_t seed, uint32_t flags,
             void *stream);

/* Bridge to the reference's dense state int64[34][R][C] (impl:16-60) -- `initial_state_override`
 * (maenv:551-553) and the facade methods use it.  status_d[b] != 0 marks a state the compact
 * layout cannot represent (never the case for states reached by play). */
int sx_import_ref_state(const sx_config *cfg, sx_state st, int64_t num_envs, const int64_t *dense_d,
                        const int8_t *player_d, uint8_t *status_d, void *stream);
int sx_export_ref_state(const sx_config *cfg, sx_state st, int64_t num_envs, int64_t *dense_d, int8_t *player_d,
                        void *stream);

/* penv.get_state_from_player_perspective (penv:101 -> impl:646-675) on the device state: the dense state as
 * viewer_d[b] sees it (player -1: player layers swapped, board rotated 180 degrees; NULL = absolute frame).
 * This is what maenv puts in the observation dict under 'internal_state' (maenv:494-495). */
int sx_export_perspective_state(const sx_config *cfg, sx_state st, int64_t num_envs, const int8_t *viewer_d,
                                int64_t *dense_d, void *stream);

/* impl:400-517 (spatial, in `player`'s frame) or impl:522-642 (1D, absolute frame): mask_d is
 * uint8 [num_envs][spatial_actions] or [num_envs][action_size].  player_d NULL = player to move. */
int sx_valid_mask(const sx_config *cfg, sx_state st, int64_t num_envs, const int8_t *player_d, int32_t format,
                  uint8_t *mask_d, void *stream);

/* maenv:447-497 without the step: mask + normalised observations for player_d (NULL = to move). */
int sx_observe(const sx_config *cfg, sx_state st, int64_t num_envs, const int8_t *player_d, sx_outputs out,
               void *stream);

/* impl:897-1045 (+ impl:835-849): applies actions_d[b] for the player to move; no rendering. */
int sx_step(const sx_config *cfg, sx_state st, int64_t num_envs, const int32_t *actions_d, int32_t action_format,
            uint32_t flags, sx_outputs out, void *stream);

/* The fused hot path, one launch per env-step: maenv.step (maenv:659-828) = action decode
 * (maenv:685-689) -> next state (impl:897) -> outcome (impl:835-849) -> [auto-reset] -> mask +
 * observations of the player to move (maenv:447-497) -> [uniform valid-action sample].
 * stats_d (optional, device, int64[8], accumulated with atomics): games finished, player +1 wins, player -1 wins,
 * invalid endings, rejected actions, actions processed, attacks, setup draws (re-draws of unplayable setups included). */
int sx_step_all(const sx_config *cfg, sx_state st, int64_t num_envs, int64_t env_base, const int32_t *actions_d,
                int32_t action_format, uint32_t flags, const uint8_t *setups_d, int32_t n_setups, uint64_t seed,
                sx_outputs out, int64_t *stats_d, void *stream);

/* Uniform draw over the valid entries of uint8 masks [num_envs][mask_len] (replaces maenv:830-834);
 * Philox4x32-10 keyed by (seed, env_base + b, step). */
int sx_sample_valid(const uint8_t *mask_d, int64_t num_envs, int32_t mask_len, int64_t env_base, uint64_t seed,
                    uint32_t step, int32_t *actions_d, void *stream);

/* Masked-logit action sampling, the step right upstream of sx_step_all in a rollout (replaces the CPU chooser
 * of examples/basic_game_loop.py:6-32 / README.md:38-65: softmax(logits + log(mask + 1e-8)) + np.random.choice).
 * Draws actions_d[b] ~ softmax(logits[b] / temperature) restricted to mask[b] != 0 (Gumbel-max, Philox4x32-10
 * keyed by (seed, env_base + b, step, entry)); invalid entries have probability exactly 0.  logits_d is
 * [num_envs][n_actions] of logits_dtype; logprob_d (optional) receives log p(action).  actions_d[b] = -1 when a
 * mask row has no valid entry. */
enum { SX_DTYPE_F32 = 0, SX_DTYPE_BF16 = 1, SX_DTYPE_F16 = 2 };
int sx_sample_logits(const void *logits_d, int32_t logits_dtype, const uint8_t *mask_d, int64_t num_envs,
                     int32_t n_actions, int64_t env_base, uint64_t seed, uint32_t step, float temperature,
                     int32_t *actions_d, float *logprob_d, void *stream);

/* The same draw taken straight from the game STATE instead of a mask: the valid entries of the player to move are
 * regenerated from the compact state (~0.3 KB per game instead of the 3.7 KB mask row of a 10x10 board), then the
 * identical Gumbel-max runs over them -- same Philox key per (seed, env_base + b, step, entry), so for the same key it
 * returns the SAME action as sx_sample_logits on that state's mask.  logits_d is [num_envs][R*C*A] in the mover's
 * frame (what a policy fed with sx_step_all's observations emits).  A finished game or stuck player yields the noop
 * entry A-1 with log-probability 0 (impl:514-515). */
int sx_sample_policy(const sx_config *cfg, sx_state st, int64_t num_envs, int64_t env_base, const void *logits_d,
                     int32_t logits_dtype, uint64_t seed, uint32_t step, float temperature, int32_t *actions_d,
                     float *logprob_d, void *stream);

/* impl:854-891 (_get_heuristic_rewards_from_move): rewards_d[b] = reward_matrix[rank of the mover's piece on the
 * start square][rank of the opponent's piece on the end square, 0 = empty] for the action game b is about to play
 * (call it BEFORE the step).  reward_matrix_d is float32 [13][13] on the device; a no-op scores 0.  Like the
 * reference, the action is assumed to be valid. */
int sx_heuristic_rewards(const sx_config *cfg, sx_state st, int64_t num_envs, const int32_t *actions_d,
                         int32_t action_format, const float *reward_matrix_d, float *rewards_d, void *stream);

/* Kernel/launch facts for the benchmark's roofline accounting. */
typedef struct {
    int32_t warps_per_block, blocks_per_sm, smem_bytes_per_block, num_sms, grid_blocks, regs_per_thread;
    int32_t background_bytes; /* shared-memory background images of a block (DESIGN.md "Rendering") */
    int32_t thread_per_game;  /* 1: the variant steps through sx_toy_kernel (one thread per game, boards of <= 16 cells) */
} sx_launch_info;
int sx_step_all_launch_info(const sx_config *cfg, uint32_t obs_mask /*1 po, 2 fo, 4 mask*/, sx_launch_info *out);

/* ---- host-buffer convenience object (the e2e path: host actions in, host tensors out) ------------
 * Owns device state + output buffers for num_envs games on the current device and pipelines
 * H2D(actions) -> sx_step_all -> D2H(outputs) in chunks over internal streams. */
typedef struct sx_host_env sx_host_env;
int sx_host_env_create(const sx_config *cfg, int64_t num_envs, int64_t env_base, uint32_t obs_mask, uint32_t flags,
                       const uint8_t *setups_host, int32_t n_setups, uint64_t seed, int32_t n_chunks,
                       sx_host_env **out);
void sx_host_env_destroy(sx_host_env *env);
/* host pointers; pinned memory recommended.  reset fills the outputs for the first player. */
int sx_host_env_reset(sx_host_env *env, sx_outputs host_out);
int sx_host_env_step(sx_host_env *env, const int32_t *actions_host, sx_outputs host_out);
/* same with an explicit action format (SX_ACTION_1D for maenv.step(..., is_spatial_index=False)) and per-call
 * flags (e.g. SX_ALLOW_OSCILLATION) instead of the ones given at creation */
int sx_host_env_step_ex(sx_host_env *env, const int32_t *actions_host, int32_t action_format, uint32_t flags,
                        sx_outputs host_out);
/* device-resident variant used for the kernel-only measurement: no host copies */
int sx_host_env_step_device(sx_host_env *env, int32_t use_sampled_actions);
int sx_host_env_sync(sx_host_env *env);
/* device pointers of the object's state (for sx_import_ref_state / sx_export_ref_state / sx_observe on it, e.g.
 * maenv's initial_state_override) and, optionally, of its device-side output buffers */
int sx_host_env_state(sx_host_env *env, sx_state *state_out, sx_outputs *device_outputs_out);

#ifdef __cplusplus
}
#endif
#endif /* STRATEGO_B200_H */
