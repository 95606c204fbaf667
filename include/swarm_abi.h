/*
 * swarm_abi.h -- C ABI of libswarm_b200.so: the B200 (sm_100a) batched implementation of the
 * gym-swarm DiscreteSwarmEnv transition.
 *
 * Reference path being replaced (all citations into /root/reference):
 *   gym_swarm/envs/swarm_discrete_env.py   ("SDE")
 *     SDE:223-267  DiscreteSwarmEnv.step           -> swarm_step
 *     SDE:197-221  DiscreteSwarmEnv.reset          -> swarm_reset (+ in-kernel auto-reset)
 *     SDE:60-80    step_all_agents, SDE:467-475 action_to_move     -> swarm_move
 *     SDE:269-292  change_position_id, SDE:246-255 re-draw loop    -> inside swarm_step
 *     SDE:90-165   Predator.__init__/closest_target/follow_target  -> swarm_predator
 *     SDE:294-380  swarm_reward                     -> swarm_pairwise (+ reward epilogue in swarm_step)
 *     SDE:382-406  set_env_parameters / set_reward_parameters      -> swarm_config fields
 *   gym_swarm/__init__.py:6-14  Gym registration (Python host side, not in this ABI)
 *
 * Conventions
 *   - plain C, no torch types: every tensor is a raw device pointer (tensor.data_ptr());
 *     the caller owns all state / tape / output memory; the library only owns small scratch.
 *   - every function returns 0 on success, a negative SWARM_E_* code on failure;
 *     swarm_last_error(h) returns a human readable message (h may be NULL for create errors).
 *   - all work is enqueued on the caller-supplied cudaStream_t (passed as void*); no hidden
 *     host threads; a handle is not thread-safe.
 *   - per-env sticky error bits (SWARM_ERR_*) are accumulated in swarm_state.err.
 *   - all randomness arrives as explicit tapes (see DESIGN.md "Tapes").
 */
#ifndef SWARM_ABI_H
#define SWARM_ABI_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SWARM_ABI_VERSION 3

/* return codes */
#define SWARM_OK 0
#define SWARM_E_INVALID (-1)   /* bad argument / unsupported shape */
#define SWARM_E_CUDA (-2)      /* CUDA runtime error (message has the cudaError string) */
#define SWARM_E_STATE (-3)     /* call order problem (state or tapes not bound) */
#define SWARM_E_NCCL (-4)      /* NCCL unavailable or failed */
#define SWARM_E_NOT_READY (-5) /* swarm_stats_end(wait = 0): the asynchronous statistics are still in flight */

/* per-env error bits (swarm_state.err) */
#define SWARM_ERR_SLAB_OVERFLOW 1u   /* collision re-draw slab (K) exhausted (SDE:249-255 has no cap) */
#define SWARM_ERR_PLACE_OVERFLOW 2u  /* placement tape exhausted during reset (SDE:207-211) */
#define SWARM_ERR_PRED_OVERFLOW 4u   /* predator placement attempts exhausted (SDE:99-105) */
#define SWARM_ERR_BAD_ACTION 8u      /* action id > 8 (reference: KeyError at SDE:239) */
#define SWARM_ERR_STEP_WHEN_DONE 16u /* stepped a finished env with auto-reset off (SDE:233-234) */

/* limits */
#define SWARM_MAX_AGENTS 16384
#define SWARM_MAX_GRID 16384
#define SWARM_FUSED_MAX_AGENTS 256   /* N <= this: fused kernels (thread per env for N <= 8 -- N <= 16 in batches of
                                        >= 65,536 envs --, lane group per env above); larger N: staged kernels */

/* swarm_config.path */
#define SWARM_PATH_AUTO 0
#define SWARM_PATH_STAGED 1
#define SWARM_PATH_FUSED 2
#define SWARM_PATH_STAGED_PAIRWISE 4 /* staged kernels, reward counts always by the tiled pairwise kernel (never the
                                       box-count path) */
#define SWARM_PATH_GROUP 3   /* fused lane-group kernel even where the thread-per-env kernel (N <= 8) would be used */

/* swarm_config.flags: kernel-selection switches for A/B measurements and for the parity tests that must reach every
 * kernel variant (read once in swarm_create; the library reads no environment variables) */
#define SWARM_FLAG_NO_FAR_BORDER 1u /* keep all four thresholds in the pair sweep (no border pass for the far ones) */
#define SWARM_FLAG_VF_DENSE 4u      /* vf_size steps of the lane-group kernels: always the dense gate sweep */
#define SWARM_FLAG_FUSED_SPLIT 8u   /* lane-group kernels: always two launches per step (transition, then reward); the default
                                       does that from 2^22 agents per handle up (instruction-cache footprint) */
#define SWARM_FLAG_FUSED_ONE_LAUNCH 16u /* lane-group kernels: always the single fused launch */
#define SWARM_FLAG_TINY_SCALAR 64u  /* thread-per-env path: the scalar kernel even for N == 8 without per-agent outputs */
#define SWARM_FLAG_NO_BOX_NEAR 32u  /* reward launch of the two-launch step: ring sweep even where box popcounts apply */

typedef struct swarm_config {
  int32_t abi_version;       /* SWARM_ABI_VERSION */
  int32_t device;            /* CUDA device ordinal */
  int32_t num_envs;          /* E: environments owned by this handle (this GPU's slice) */
  int32_t num_agents;        /* N   (SDE:184) */
  int32_t grid_size;         /* S   (SDE:185) */
  int32_t attraction_thresh; /* ceil() of SDE:193 -- integer distances make ceil exact */
  int32_t repulsion_thresh;  /* ceil() of SDE:194 */
  float predator_eat_rew;    /* SDE:195 */
  int32_t slab_len;          /* K : re-draw actions available per env-step */
  int32_t place_len;         /* PL: int16 values per env per reset slot (2N + 2*KR) */
  int32_t pred_attempts;     /* KP: predator placement attempts per reset slot */
  int32_t reset_slots;       /* R : reset tape slots; n-th reset of an env uses slot n % R */
  int32_t auto_reset;        /* 1: finished envs reset in-kernel; 0: they freeze */
  int32_t track_stats;       /* 1: maintain episode statistics (needs the stats pointers) */
  int32_t path;              /* SWARM_PATH_* */
  uint32_t flags;            /* SWARM_FLAG_* */
  int32_t tiny16_min_envs;   /* 9..16 agents: thread-per-env kernel from this many envs up (0: default 65,536) */
  uint32_t env_offset;       /* global index of this handle's env 0 (its shard's `lo`): the counter-based (Philox) tapes
                                are keyed by env + env_offset, so a sharded job draws what the unsharded one draws */
  int32_t reserved;          /* 0 */
} swarm_config;

/* Persistent device state, all caller-allocated. SoA; positions are int16. */
typedef struct swarm_state {
  int16_t* x;            /* [E,N] agent x (SDE:214 current_state[i][0]) */
  int16_t* y;            /* [E,N] agent y */
  int16_t* px;           /* [E]   predator x (SDE:92 Predator.current_state) */
  int16_t* py;           /* [E]   predator y */
  int32_t* target;       /* [E]   Predator.current_target (SDE:109,140) */
  uint8_t* pred_orient;  /* [E]   Predator.orientation (SDE:107,163) */
  uint8_t* frozen;       /* [E]   1 = episode finished and not reset (auto_reset == 0) */
  uint32_t* err;         /* [E]   sticky SWARM_ERR_* bits */
  uint32_t* reset_count; /* [E]   number of resets performed (selects the reset slot) */
  int16_t* tcell;        /* [E,2] (opt) cell of Predator.current_target: the agent position follow_target reads at the
                            next step (SDE:142).  State of the STAGED path only (N > 256): its predator and reset kernels
                            maintain it, and with it the predator kernel needs no dependent random read of the old
                            positions; the fused kernels hold the whole env in registers and ignore it.  NULL: not kept.
                            A caller that overwrites x / y / target itself must refresh it (or pass NULL) */
  /* episode statistics (track_stats == 1, else may be NULL) */
  uint32_t* ep_rec;      /* [E,4] running episode, one 16-byte record: word 0 = length (u32), words 1..3 = f32 bits of
                            the running attraction, repulsion, alignment return (the survival term is non-zero only on
                            the terminal step, so a running episode's survival return is always 0) */
  uint32_t* fin_count;   /* [E]   finished episodes */
  uint32_t* fin_len;     /* [E]   sum of finished episode lengths */
  float* fin_ret;        /* [E,4] sum of finished episode returns per component */
} swarm_state;

/* Per-step inputs. */
typedef struct swarm_step_in {
  const uint8_t* actions;  /* [E,N] action ids 0..8 (SDE:467-475) */
  const uint8_t* retarget; /* [E]   1 iff the reference's np.random.random() < 0.1 (SDE:138-139) */
  const uint8_t* slab;     /* [E,K] re-draw actions 0..7 in consumption order (SDE:250) */
                           /* retarget == slab == NULL: both tapes are drawn in the kernels by Philox4x32-10 from
                              (rng_seed, rng_step): counter-based, no tape memory (DESIGN.md "Tapes") */
  double w_attraction;     /* reward_type["attraction"] (bool or float: the reward curriculum) */
  double w_repulsion;      /* reward_type["repulsion"] */
  double w_alignment;      /* reward_type["alignment"] */
  int32_t vf_size;         /* floor(reward_type["vf_size"]) or -1 for None (SDE:349-355); a NEGATIVE Python vf_size clears
                              every pair (alignment term 0): the caller passes -1 together with w_alignment = 0 */
  uint32_t rng_step;       /* device tapes: index of this step (the caller increments it) */
  uint64_t rng_seed;       /* device tapes: key of the step streams */
} swarm_step_in;

/* Per-step outputs; pointers marked (opt) may be NULL. */
typedef struct swarm_step_out {
  uint8_t* done;         /* [E]   done flag of this step (SDE:263) */
  int32_t* eaten;        /* [E]   id of the agent the predator landed on, -1 if none (SDE:323) */
  float* reward;         /* [E,5] global reward: survival, attraction, repulsion, alignment, sum */
  int32_t* counts;       /* [E,2] integer rew_rep (SDE:338), rew_attr (SDE:343); 0 on termination */
  int16_t* pred_prev;    /* [E,2] info["predator_state"] (SDE:258,264) */
  int16_t* pred_next;    /* [E,2] info["predator_next_state"] (SDE:266) */
  uint8_t* step_orient;  /* [E]   info["predator_orientation"] (SDE:265) */
  int32_t* step_target;  /* [E]   predator target after this step, before any auto-reset */
  int32_t* consumed;     /* [E]   number of re-draw actions consumed */
  int16_t* term_x;       /* [E,N] (opt) positions produced by the step, written where done == 1 */
  int16_t* term_y;       /* [E,N] (opt) */
  float* agent_reward;   /* [E,N,4] (opt) per-agent attraction, repulsion, alignment, sum (SDE:371-379);
                            survival is predator_eat_rew for agent `eaten`, 0 otherwise */
  int32_t* agent_counts; /* [E,N,2] (opt) integer rew_rep_i (SDE:372), rew_attr_i (SDE:373) */
  uint8_t* eff_action;   /* [E,N] (opt) action id each agent finally moved with (after re-draws) */
  int16_t* agent_terms;  /* [E,N,4] (opt, N <= 4096) compact per-agent credit: the integers the per-agent floats are exact
                            functions of -- (rew_rep_i, rew_attr_i, U_i, Q_i) with
                              repulsion_i = w_rep * rew_rep_i / (N(N-1)),  attraction_i = w_att * rew_attr_i / (N(N-1)),
                              alignment_i = -w_ali * (0.25 * U_i - (sqrt(2)/4) * Q_i) / (N(N-1))      (SDE:371-379);
                            all four are 0 on a terminal step.  Half the bytes of agent_reward for callers that read
                            rewards over PCIe; either, both or neither may be bound */
} swarm_step_out;

typedef struct swarm_ctx* swarm_handle;

int swarm_create(const swarm_config* cfg, swarm_handle* out);
int swarm_destroy(swarm_handle h);
const char* swarm_last_error(swarm_handle h);
int swarm_abi_version(void);

int swarm_bind_state(swarm_handle h, const swarm_state* st);
/* place: int16[R,E,PL]  ptape: int16[R,E,KP,2] */
int swarm_bind_reset_tapes(swarm_handle h, const int16_t* place, const int16_t* ptape);

/* Reset tapes drawn in the kernels instead (Philox4x32-10 keyed by seed, counter = (env, reset index, ...)): unbinds
 * any bound reset tapes; the n-th reset of an env then uses stream index n (no wrap-around at reset_slots). */
int swarm_set_device_reset_tapes(swarm_handle h, uint64_t seed);

/* SDE:197-221 for every env (reset_count := 0, slot 0). */
int swarm_reset(swarm_handle h, void* stream);
/* SDE:223-267 for every env, plus auto-reset / statistics as configured. */
int swarm_step(swarm_handle h, const swarm_step_in* in, const swarm_step_out* out, void* stream);
/* Name of the kernel path swarm_step uses for this handle ("tiny<8>", "fused<LPE,APL>" or "staged"). */
const char* swarm_path_name(swarm_handle h);
/* Number of kernel launches the last swarm_step / swarm_reset enqueued. */
int swarm_last_launch_count(swarm_handle h);

/* ---- standalone kernels (roofline runs; they are also the stages of the staged path) ---- */

/* SDE:60-80 + SDE:467-475: out = wrap(pos + move[action]); no collision handling. HBM bound. */
int swarm_move(int32_t E, int32_t N, int32_t S, const int16_t* x, const int16_t* y, const uint8_t* actions,
               int16_t* out_x, int16_t* out_y, void* stream);
/* SDE:133-165 + SDE:111-131 + the termination test SDE:307-309: one warp per env, Chebyshev argmin by
 * warp shuffles over the agents' OLD positions, pursuit, wrap; done/eaten against the NEW positions. */
int swarm_predator(int32_t E, int32_t N, int32_t S, const int16_t* x_old, const int16_t* y_old,
                   const int16_t* x_new, const int16_t* y_new, const uint8_t* retarget, int16_t* px, int16_t* py,
                   int32_t* target, uint8_t* pred_orient, int16_t* pred_prev, uint8_t* done, int32_t* eaten,
                   int16_t* tcell /* [E,2] in/out (opt): swarm_state.tcell; NULL: the target's old cell is read from x_old/y_old */,
                   void* stream);
/* SDE:330-343 integer part: per-agent un-weighted rew_rep_i / rew_attr_i from a tiled O(N^2) pass.
 * out_counts: int32[E,N,2]; skip (may be NULL): uint8[E], envs with skip != 0 are not computed. */
int swarm_pairwise(int32_t E, int32_t N, int32_t S, int32_t attraction_thresh, int32_t repulsion_thresh,
                   const int16_t* x, const int16_t* y, const uint8_t* skip, int32_t* out_counts, void* scratch,
                   uint64_t scratch_bytes, void* stream);
/* Device scratch swarm_pairwise needs for (E, N): the caller allocates it (any 256-byte aligned device block) and may
 * reuse it between calls on one stream; calls with different scratch blocks may overlap freely. */
uint64_t swarm_pairwise_scratch_bytes(int32_t E, int32_t N);
/* SDE:197-221: reset the envs whose mask byte is non-zero (all if mask == NULL) from their next reset slot. */
int swarm_autoreset(swarm_handle h, const uint8_t* mask, void* stream);

/* Visual-field observations (the policy-side output after a step; README figure "Visual Field 5x5"): for every agent
 * the egocentric (2*vf+1)^2 window, out uint8[E,N,2*vf+1,2*vf+1] indexed [dy+vf][dx+vf]: 0 empty / outside the grid
 * (plain Chebyshev distance, as the vf gate of SDE:349-355), 1 the agent itself, 2 another agent, 3 the predator.
 * N <= 4096, 0 <= vf <= 7.  The generating code is not in the reference tree: the codes are this library's definition. */
int swarm_observe_vf(int32_t E, int32_t N, int32_t S, int32_t vf, const int16_t* x, const int16_t* y, const int16_t* px,
                     const int16_t* py, uint8_t* out, void* stream);

/* ---- episode statistics ---- */
/* out[8] = episodes, sum length, sum return, sum survival, sum attraction, sum repulsion, sum alignment,
 * env-steps of running episodes.  Deterministic two-stage reduction over the envs (multi-CTA partials, fixed order).
 *
 * Asynchronous form (the one a step loop uses, e.g. every 64 steps): swarm_stats_begin enqueues the reduction on
 * `stream` (so it sees exactly the steps enqueued before it), then -- on the handle's own side stream, behind an event,
 * off the step's critical path -- the optional NCCL sum all-reduce of the 8 doubles ON THE DEVICE BUFFER and the copy
 * into a pinned host slot.  It never synchronises the host.  swarm_stats_end returns the oldest result begun and not
 * yet collected: SWARM_E_NOT_READY if wait == 0 and it is still in flight.  Up to 4 results may be in flight. */
int swarm_stats_begin(swarm_handle h, int32_t allreduce, void* stream);
int swarm_stats_end(swarm_handle h, double out[8], int32_t wait);
/* Synchronous conveniences: begin + end(wait = 1). */
int swarm_stats_read(swarm_handle h, double out[8], void* stream);
int swarm_stats_allreduce(swarm_handle h, double out[8], void* stream);
/* NCCL (dlopen'ed libnccl.so.2): the library bootstraps its own communicator from a unique id the host broadcasts
 * (swarm_nccl_init), or adopts one the caller already owns (swarm_nccl_adopt: comm is an ncclComm_t; not destroyed by
 * swarm_destroy). */
int swarm_nccl_unique_id(uint8_t out_id[128]);
int swarm_nccl_init(swarm_handle h, const uint8_t id[128], int32_t nranks, int32_t rank);
int swarm_nccl_adopt(swarm_handle h, void* comm);

/* ---- measurement helpers ---- */
/* INT32-pipe micro-benchmarks (independent dependent-chain kernels); returns lane-ops/s in *out.
 * kind: 0 IADD3, 1 IMNMX, 2 IADD3+IMAD mix, 3 VIADDMNMX.S16x2 (DPX), 4 VIADDMNMX.S16x2 + IMAD mix */
int swarm_int32_peak(int32_t device, int32_t kind, double* out_lane_ops_per_s, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SWARM_ABI_H */
