/* oracle/qhg_oracle.h -- C API of the CPU restatement of QHG4's per-step agent update.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline leg may load this library; the product (qhg4_b200/) never does.
 *
 * The restatement is single-threaded and follows the reference structurally: a slot array
 * with holes, one full pass over the slots per action in priority order, queued
 * births/deaths/moves applied in finalizeStep (core/SPopulation.cpp:394-477,554-577,596-724).
 * It runs in two modes:
 *   QOR_MODE_WELL     the reference's own random streams and slot recycling for ONE OpenMP
 *                     thread; pinned bit-exactly against oracle/_ref (tests/test_oracle_vs_ref.py)
 *   QOR_MODE_COUNTER  identical action logic, but every draw comes from the counter-based
 *                     generator the CUDA path uses (Philox4x32-10 keyed by agent ID, step and
 *                     stream), pairing is by random-key rank and newborn IDs by (cell, mother ID)
 *                     rank, so results do not depend on slot order.  The CUDA path is compared
 *                     bit-exactly against this mode.
 */
#ifndef QHG_ORACLE_H
#define QHG_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define QOR_MODE_WELL    0
#define QOR_MODE_COUNTER 1

typedef struct qor_pop qor_pop;

qor_pop *qor_create(const char *pop_class, int n_cells, int max_neigh, int mode);
void     qor_destroy(qor_pop *p);
int      qor_set_cells(qor_pop *p, const int32_t *nbr, const int32_t *global_id);
int      qor_set_env_array(qor_pop *p, const char *name, const double *v, int64_t n);
int      qor_set_env_delta(qor_pop *p, const char *name, const double *delta, int64_t n);  /* core/AutoInterpolator.cpp:461-483 */
int      qor_interpolate_env(qor_pop *p, int steps);
int      qor_get_env_array(qor_pop *p, const char *name, double *out);
int      qor_set_attribute(qor_pop *p, const char *name, double v);
int      qor_set_attribute_str(qor_pop *p, const char *name, const char *v);
int      qor_set_prio(qor_pop *p, const char *action, int prio);
int      qor_enable_action(qor_pop *p, const char *action, int enabled);
int      qor_set_seed(qor_pop *p, const uint32_t *state16);
int      qor_add_agents(qor_pop *p, int64_t n, const int32_t *cell, const int64_t *id, const float *birth,
                        const uint8_t *gender, const float *age, const float *last_birth, const uint32_t *life);
int      qor_pre_loop(qor_pop *p);
int      qor_initialize_step(qor_pop *p, float t);
int      qor_do_actions(qor_pop *p, unsigned prio, float t);
int      qor_finalize_step(qor_pop *p);
int      qor_step(qor_pop *p, float t);
int      qor_update_event(qor_pop *p, int event_id, float t);
int      qor_flush_events(qor_pop *p, float t);

int64_t  qor_get_num_agents_effective(qor_pop *p);
int      qor_get_num_agents_array(qor_pop *p, uint64_t *out);
int64_t  qor_get_agents(qor_pop *p, int64_t cap, int32_t *cell, int64_t *id, float *birth, uint8_t *gender,
                        float *age, float *last_birth, uint32_t *life, int64_t *mate_id, int32_t *slot);
int      qor_get_env_weights(qor_pop *p, double *out);             /* nCells*(maxNeigh+1) */
int      qor_get_birth_death_probs(qor_pop *p, double *b, double *d);
int      qor_get_capacities(qor_pop *p, double *out);
int      qor_get_move_stats(qor_pop *p, int32_t *hops, double *dist, double *time);  /* actions/MoveStats.cpp: per-cell arrays */
int      qor_atan_death_prob(qor_pop *p, int n, const float *age, double *out);
int      qor_get_step_stats(qor_pop *p, uint64_t *births, uint64_t *deaths, uint64_t *moves);

int      qor_set_navigation(qor_pop *p, int n_ports, const int32_t *port_cell, const int32_t *port_ptr, const int32_t *dest_cell,
                            const double *dist, int n_bridges, const int32_t *bridges);

/* genomes (actions/Genetics.cpp, genes/BitGeneUtils.cpp) */
int      qor_set_genomes(qor_pop *p, int64_t n, const uint64_t *g);
int64_t  qor_get_genomes(qor_pop *p, int64_t cap, uint64_t *g, int32_t *num_babies);
int      qor_bit_crossover(const uint32_t *state16, const uint64_t *in, int genome_size, int n_cross, uint64_t *out);
int      qor_bit_freereco(const uint32_t *state16, const uint64_t *in, int n_blocks, uint64_t *out);
int      qor_bit_mutate(const uint32_t *state16, uint64_t *genome, int n_bits, int n_mut);
int      qor_gene2_crossover(const uint32_t *state16, const uint64_t *in, int genome_size, int n_cross, uint64_t *out);  /* genes/GeneUtils.cpp */
int      qor_gene2_freereco(const uint32_t *state16, const uint64_t *in, int n_blocks, uint64_t *out);
int      qor_gene2_mutate(const uint32_t *state16, uint64_t *genome, int n_nucs, int n_mut);
int      qor_set_genetics_well(qor_pop *p, const uint32_t *state16, uint32_t index);
int      qor_binomial_table(double prob, int n, double eps, int cap, double *out);
int      qor_binomial_get_n(double prob, int n, double eps, double r);

/* sharded runs: the protocol of the CUDA path's multi-GPU mode (counter mode) */
int64_t  qor_get_pending_births(qor_pop *p);
int      qor_set_birth_id_offset(qor_pop *p, int64_t offset, int64_t total);
int64_t  qor_extract_foreign(qor_pop *p, int c0, int c1, int64_t cap, int32_t *cell, int64_t *id, float *birth, uint8_t *gender,
                             float *age, float *last_birth, uint32_t *life);
int      qor_recount(qor_pop *p);
int64_t  qor_get_max_id(qor_pop *p);
int      qor_set_max_id(qor_pop *p, int64_t v);

/* stand-alone pieces */
void     qor_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
int      qor_well_sequence(const uint32_t *state16, int n, uint32_t *out);
int      qor_polyline_eval(const char *def, int n, const double *x, double *out, int float_cast);
#ifdef __cplusplus
}
#endif
#endif
