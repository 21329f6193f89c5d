/*
 * mindrl_b200_aot.h -- the experience data path behind MindSpore's "aot" custom-operator ABI.
 *
 * The reference mounts native operators only through
 *     ops.Custom("<lib>.so:<Func>", out_shapes, out_dtypes, "aot", reg_info=...)
 * (mindspore_rl/utils/mcts/mcts.py:247-254), which binds the two C symbols declared per operator below:
 * the operator itself and <Func>Init (mindspore_rl/utils/mcts/cpu/cpu_mcts_ops.cc:38-53).
 *   params  inputs first, then outputs; device pointers on GPU
 *   ndims / shapes / dtypes  one entry per parameter; dtypes are MindSpore's names ("float32", "int64", "bool")
 *   stream  cudaStream_t
 *   extra   AotExtra * (mindspore_rl/utils/mcts/custom_aot_extra.h:30-105; our restatement of the interface is
 *           mindrl_b200/csrc/aot_extra.h): attributes registered with CustomRegOp.attr, per-operator data
 *   return  0 = ok, 2 = failure (kErrorCode, cpu_mcts_ops.cc:25); text through mrl_last_error()
 * libmindrl_b200.so exports these next to the handle-based C ABI of mindrl_b200.h; INTEGRATION.md shows the
 * Python side that mounts them.  Parameter lists (inputs | outputs) and attributes:
 *
 *   BufferAppend                 buffer[0..n) exp[0..n) count head | count_out int32[1]
 *                                exp[i] is one row or N rows of buffer[i]; count, head int32[1], updated in place
 *                                replaces P.BufferAppend, mindspore_rl/core/uniform_replay_buffer.py:82, 114
 *   BufferGetItem                buffer[0..n) count head index(int32|int64 [1]) | row[0..n)      (:86, 128)
 *   BufferSample                 buffer[0..n) count head | batch[0..n)   attrs: seed, unique=false  (:84-89, 139)
 *   PriorityReplayBufferCreate   | handle int64[1]   attrs: capacity alpha shapes(int 2-D) dtypes(str, comma
 *                                separated) seed0 seed1     mindspore_rl/core/priority_replay_buffer.py:62-66
 *   PriorityReplayBufferPush     transition[0..n) | handle int64[1]          attrs: handle            (:67, 90)
 *   PriorityReplayBufferSample   beta float32[1] | indices int64[B] weights float32[B] batch[0..n)
 *                                attrs: handle                                                      (:68-70, 106)
 *   PriorityReplayBufferUpdate   indices int64[B] priorities float32[B] | handle int64[1]  attrs: handle (:71, 120)
 *   PriorityReplayBufferDestroy  | handle int64[1]                             attrs: handle           (:138)
 *   DiscountedReturn             reward[T,B,..] done bool[T,B] last_state_value[B,..] | returns[T,B,..]
 *                                attrs: gamma                 mindspore_rl/utils/discounted_return.py:71, 82
 *   GaeScan                      reward value last_value done | advantage returns
 *                                attrs: gamma lambda layout(0 = [T,B], 1 = [B,T]) normalize(bool)
 *                                the in-learner loops ppo.py:334-368, mappo.py:478-497 as one operator
 */
#ifndef MINDRL_B200_AOT_H_
#define MINDRL_B200_AOT_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MRL_AOT_OP(name)                                                                                   \
  int name(int nparam, void **params, int *ndims, int64_t **shapes, const char **dtypes, void *stream,    \
           void *extra);                                                                                   \
  int name##Init(int *ndims, int64_t **shapes, const char **dtypes, void *extra /* AotExtra * */)

MRL_AOT_OP(BufferAppend);
MRL_AOT_OP(BufferGetItem);
MRL_AOT_OP(BufferSample);
MRL_AOT_OP(PriorityReplayBufferCreate);
MRL_AOT_OP(PriorityReplayBufferPush);
MRL_AOT_OP(PriorityReplayBufferSample);
MRL_AOT_OP(PriorityReplayBufferUpdate);
MRL_AOT_OP(PriorityReplayBufferDestroy);
MRL_AOT_OP(DiscountedReturn);
MRL_AOT_OP(GaeScan);

/* A concrete AotExtra for hosts that are not MindSpore (the parity tests, other frameworks): create one,
 * set the operator's attributes, pass it as `extra` to <Func>Init and <Func>; destroy frees the operator's
 * data too. */
void *mrl_aot_extra_create(void);
void mrl_aot_extra_destroy(void *extra);
int mrl_aot_extra_set_bool(void *extra, const char *name, int32_t value);
int mrl_aot_extra_set_int(void *extra, const char *name, int64_t value);
int mrl_aot_extra_set_float(void *extra, const char *name, float value);
int mrl_aot_extra_set_str(void *extra, const char *name, const char *value);
/* rows of a 2-D integer attribute, flattened: row r has lens[r] entries */
int mrl_aot_extra_set_int2d(void *extra, const char *name, const int64_t *flat, const int32_t *lens,
                            int32_t nrows);

#ifdef __cplusplus
}
#endif
#endif /* MINDRL_B200_AOT_H_ */
