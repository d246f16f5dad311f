/*
 * rs_host.h -- host-preparation pieces of libresynthesizer_b200.so exported so that parity tests can compare
 * them with the oracle array by array.  They run on the CPU only (no CUDA call).
 */
#ifndef RS_HOST_H
#define RS_HOST_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
/* replaces quantizeMetricFuncs (lib/matchWeighting.h:194-204): tables indexed 256 + signed difference */
void rs_host_metric_tables(double sensitivity, double map_weight, uint16_t *color512, uint32_t *map512);
/* replaces prepareSortedOffsets (lib/engine.c:465-497); writes min(cap, n) x,y pairs, returns n */
uint32_t rs_host_sorted_offsets(int tw, int th, int cw, int ch, int32_t *xy, uint32_t cap);
/* replaces orderTargetPoints (lib/orderTarget.h:268-343) over n x,y pairs given in row-major scan order */
int rs_host_order_targets(int match_context_type, int32_t *xy, uint32_t n, uint32_t seed);
/* replaces prepare_repetition_parameters (lib/passes.h:67-93) */
/* `count` successive g_rand_int_range(0, n) draws of the GRand stream seeded with `seed`: directly (via_raw_stream = 0)
 * or through the early-started raw-word producer the engine uses (1).  Must be identical. */
void rs_host_draws(uint32_t seed, uint32_t n, uint32_t count, uint32_t *out, int via_raw_stream);
uint32_t rs_host_pass_schedule(uint32_t n_targets, uint32_t *ends6);
/* MT19937 jump-ahead (csrc/host_prep.h: mt_jump_poly): the positions of the set bits of z^(q * jump_words) mod phi(z),
 * ascending, into idx (room for 19937 entries); returns their number.  The state of the generator after q * jump_words
 * words is the XOR of the windows of the untempered stream that start at those positions. */
uint32_t rs_host_mt_jump_poly(uint32_t q, uint32_t jump_words, uint16_t *idx);
#ifdef __cplusplus
}
#endif
#endif
