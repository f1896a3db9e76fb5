/* qhg_b200.h -- C ABI of the B200-native per-step agent update for QHG4 populations.
 *
 * One `qhgb_pop` replaces one `SPopulation<T>` object of the reference behind its `PopBase`
 * interface (core/PopBase.h:16-121): the host keeps calling initializeStep / doActions /
 * finalizeStep once per step from ONE thread; agent storage, every action and the
 * birth/death/move bookkeeping run as sm_100a CUDA kernels on device-resident
 * structure-of-arrays state.  Plain pointers and sizes only; no C++ or torch types.
 *
 * Conventions follow the reference's: every call returns int 0 on success and non-zero
 * (normally -1) on error, results may be summed by the caller (core/PopLooper.cpp:166-202),
 * no exception crosses the boundary; `qhgb_last_error()` gives the text the reference would
 * have printed.  All paths cited below are relative to /root/reference/QHG4/.
 *
 * There is NO CPU fallback: without a CUDA device `qhgb_create` fails.
 */
#ifndef QHG_B200_H
#define QHG_B200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct qhgb_pop qhgb_pop;

/* event ids that cross the boundary (utils_qhg/EventConsts.h:14-36) */
#define QHGB_EVENT_ID_GEO      2
#define QHGB_EVENT_ID_CLIMATE  3
#define QHGB_EVENT_ID_VEG      4
#define QHGB_EVENT_ID_NAV      5
#define QHGB_EVENT_ID_FLUSH   20

/* life states in uploaded / downloaded agent records (core/SPopulation.h:70-74) */
#define QHGB_LIFE_STATE_DEAD     0u
#define QHGB_LIFE_STATE_ALIVE    1u
#define QHGB_LIFE_STATE_FERTILE  5u

/* per-step totals, the numbers the reference prints from recycleDeadSpaceNew / performMoves
 * (core/SPopulation.cpp:605-607,1087) plus the live count */
typedef struct qhgb_step_stats {
    int64_t num_agents;   /* live agents after the step = getNumAgentsEffective(), core/SPopulation.h:148 */
    int64_t births;
    int64_t deaths;
    int64_t moves;
    int64_t next_id;      /* next agent id to be handed out (replaces IDGen, core/IDGen.h:17-36) */
    int64_t steps_done;
} qhgb_step_stats;

/* ---- life cycle --------------------------------------------------------------------------------
 * qhgb_create replaces the plugin's createPop(...) -> `new <Pop>(pCG, pPopFinder, iLayerSize, apIDG,
 * aulState, aiSeeds)` (dynpops/WrapperTemplate.cpp.tmp:24-30, core/SPopulation.cpp:45-87).
 * `pop_class` selects the action set and wiring of a shipped population class, e.g.
 * "tut_EnvironAltPop" (populations/tut_EnvironAltPop.cpp:24-53).  `capacity_hint` = expected maximum
 * number of live agents (0: grow on demand); it replaces iLayerSize. */
int  qhgb_create(const char *pop_class, int device, int n_cells, int max_neigh, int64_t capacity_hint, qhgb_pop **out);
int  qhgb_destroy(qhgb_pop *p);                     /* ~SPopulation, core/SPopulation.cpp:93-134 */
const char *qhgb_last_error(void);
const char *qhgb_version(void);

/* ---- grid and environment (what the actions read through SCellGrid* / Geography*) -----------------
 * nbr: n_cells*max_neigh cell indices, -1 padded  = SCell::m_aNeighbors (core/SCell.h:9-13);
 * global_id: n_cells ids or NULL for id == index   = SCell::m_iGlobalID. */
int  qhgb_set_cells(qhgb_pop *p, const int32_t *nbr, const int32_t *global_id);
/* name: "Altitude" (Geography::m_adAltitude), "Ice" (m_abIce, non-zero = ice), "Water", "Coastal", "Latitude",
 * "Longitude" (core/Geography.h:31-39), "NPP" ... one double per cell.  Calling it again after preLoop is
 * how a changed array reaches the device; follow it with qhgb_update_event as the host does
 * (app/Simulator.cpp:668-753). */
int  qhgb_set_env_array(qhgb_pop *p, const char *name, const double *values, int64_t n);

/* Interpolated environment (core/AutoInterpolator.cpp:461-483, called from the simulator's loop app/Simulator.cpp:338-352,
 * 356-367 between checkEvents and flushEvents): between two dated environment files every target array grows by a fixed
 * per-step difference array.  qhgb_set_env_delta hands over the difference array of one target (m_mDiff[name]; NULL
 * removes it) -- it stays on the device; qhgb_interpolate_env(steps) = AutoInterpolator::interpolate(iSteps): every target
 * array += steps * its difference array, on the device, no host traffic.  The host then delivers the interpolator's events
 * (qhgb_update_event for EVENT_ID_GEO / _CLIMATE / _VEG) and qhgb_flush_events exactly as after a reloaded array. */
int  qhgb_set_env_delta(qhgb_pop *p, const char *name, const double *delta, int64_t n);
int  qhgb_interpolate_env(qhgb_pop *p, int steps);
/* the current values of an environment array on the device (after interpolation the host's Geography / Climate / Vegetation
 * copy is stale; a host that writes them out -- e.g. the "write env" event, app/Simulator.cpp -- reads them back here) */
int  qhgb_get_env_array(qhgb_pop *p, const char *name, double *out);

/* the Navigation group (sea-ways) that Navigate reads through SCellGrid::m_pNavigation (core/Navigation.h:13-37,
 * io/NavGroupReader.cpp:98-160): n_ports origin cells; port p jumps to dest_cell[port_ptr[p] .. port_ptr[p+1]) over
 * the distances dist[] (same indexing); bridges = n_bridges pairs of cells.  Navigate builds its jump tables from it
 * at preLoop and again after EVENT_ID_NAV / EVENT_ID_GEO + flush (actions/Navigate.cpp:79-144). */
int  qhgb_set_navigation(qhgb_pop *p, int n_ports, const int32_t *port_cell, const int32_t *port_ptr, const int32_t *dest_cell,
                         const double *dist, int n_bridges, const int32_t *bridges);

/* ---- parameters --------------------------------------------------------------------------------
 * Attribute names are the reference's XML / QDF attribute names, e.g. "ATanDeath_max_age",
 * "Verhulst_K", "WeightedMove_prob" (actions/ATanDeath.h:9-12, actions/Verhulst.h:9-13, ...).
 * qhgb_set_attribute     = Action<T>::tryGetAttributes / modifyAttributes for numeric values
 *                          (actions/Action.h:27-58, core/SPopulation.cpp "modifyAttributes");
 * qhgb_set_attribute_str = the same for string-valued ones (poly-lines such as "AltCapPref",
 *                          utils/PolyLine.cpp:92-127);
 * qhgb_set_prio          = one <prio name= value=> entry -> Prioritizer::setPrio (core/Prioritizer.cpp:19-30);
 *                          action names are "Name" or "Name[id]" (actions/Action.cpp:11-19);
 * qhgb_enable_action     = PopBase::enableAction / disableAction (core/PopBase.h:22-24);
 * qhgb_set_seed          = the 16-word WELL state argument `aulState` of the constructor. */
int  qhgb_set_attribute(qhgb_pop *p, const char *name, double value);
int  qhgb_set_attribute_str(qhgb_pop *p, const char *name, const char *value);
int  qhgb_set_prio(qhgb_pop *p, const char *action_name, int prio);
int  qhgb_enable_action(qhgb_pop *p, const char *action_name, int enabled);
int  qhgb_set_seed(qhgb_pop *p, const uint32_t *state16);

/* ---- agents ------------------------------------------------------------------------------------
 * Structure-of-arrays view of the reference's agent record (core/SPopulation.h:43-50 +
 * populations/tut_EnvironAltPop.h:16-21).  qhgb_add_agents = SPopulation::addAgent /
 * readAgentDataQDF (core/SPopulation.cpp:1149-1166,1689-1741); agents with life == 0 are skipped.
 * qhgb_get_agents = what preWrite + writeAgentDataQDFSafe hand to the file layer
 * (core/SPopulation.cpp:1465-1568): live agents, grouped by cell; any pointer may be NULL.
 * It returns the number of live agents (and fills at most `cap` of them), or -1. */
int     qhgb_add_agents(qhgb_pop *p, int64_t n, const int32_t *cell, const int64_t *id, const float *birth_time,
                        const uint8_t *gender, const float *age, const float *last_birth, const uint32_t *life_state);
int64_t qhgb_get_agents(qhgb_pop *p, int64_t cap, int32_t *cell, int32_t *cell_id, int64_t *id, float *birth_time,
                        uint8_t *gender, float *age, float *last_birth, uint32_t *life_state, int64_t *mate_id);

/* genomes of populations with Genetics<.., BitGeneUtils> (actions/Genetics.cpp:196-267,285-337; genes/BitGeneUtils.cpp):
 * 2 strands x ceil(Genetics_genome_size/64) 64-bit words per agent, 1 bit per nucleotide.
 * qhgb_set_genomes = what readAgentDataQDF + SequenceIOUtils load beside the agents (actions/Genetics.cpp "readAdditionalDataQDF"):
 *                    one row per agent added so far, in the order of qhgb_add_agents; before qhgb_pre_loop; rows never set are zero.
 * qhgb_get_genomes = what writeAdditionalDataQDF writes: one row per live agent in the order of qhgb_get_agents, plus
 *                    m_iNumBabies (populations/OoANavGenPop.h:21-27).  Returns the number of live agents, or -1. */
int     qhgb_set_genomes(qhgb_pop *p, int64_t n, const uint64_t *genomes);
int64_t qhgb_get_genomes(qhgb_pop *p, int64_t cap, uint64_t *genomes, int32_t *num_babies);

/* ---- the step (PopBase virtuals called by PopLooper::doStep, core/PopLooper.cpp:166-202) -----------
 * qhgb_pre_loop        = PopBase::preLoop (core/SPopulation.cpp:273-292): uploads are sealed, agents are binned
 *                        per cell, per-cell counts and derived parameters are computed.
 * qhgb_initialize_step = initializeStep(t) (core/SPopulation.cpp:394-417): Verhulst b/d from last step's counts,
 *                        pairing, cell weights if the environment changed.
 * qhgb_do_actions      = doActions(prio, t) (core/SPopulation.cpp:554-577), once per priority level.  The device
 *                        runs all levels that were requested during the step as ONE fused pass when
 *                        qhgb_finalize_step is called; order and visibility rules are the reference's.
 * qhgb_finalize_step   = finalizeStep() (core/SPopulation.cpp:439-477): births, deaths, moves, re-binning, counts.
 * qhgb_step            = the three above for all of this population's levels (PopLooper::doStep for one pop). */
int  qhgb_pre_loop(qhgb_pop *p);
int  qhgb_initialize_step(qhgb_pop *p, float t);
int  qhgb_do_actions(qhgb_pop *p, unsigned prio, float t);
int  qhgb_finalize_step(qhgb_pop *p);
int  qhgb_step(qhgb_pop *p, float t);
/* n steps at t0, t0+1, ... without returning to the host in between (the bench's device-resident loop) */
int  qhgb_run(qhgb_pop *p, float t0, int n_steps);
/* sums over all completed steps since preLoop: live agents at step start (the "agent-steps" of a throughput figure), agents
 * sent to / received from other ranks; any pointer may be NULL */
int  qhgb_get_run_totals(qhgb_pop *p, int64_t *agent_steps, int64_t *sent, int64_t *received);
int  qhgb_synchronize(qhgb_pop *p);

/* updateEvent(id, data, t) / flushEvents(t) (populations/tut_EnvironAltPop.cpp:93-127) */
int  qhgb_update_event(qhgb_pop *p, int event_id, float t);
int  qhgb_flush_events(qhgb_pop *p, float t);

/* ---- state read back by the host -------------------------------------------------------------------
 * getNumAgentsEffective / getNumAgentsTotal (core/SPopulation.h:147-148), getNumAgentsArray / getNumAgents(cell)
 * (:145-146; one ulong per cell). */
int64_t qhgb_get_num_agents_effective(qhgb_pop *p);
int     qhgb_get_num_agents_array(qhgb_pop *p, uint64_t *out);
/* the same for the cells [cell_begin, cell_end) only, out[0] = cell_begin's count: what a rank of a sharded run reads
 * back every step (its own cell range; the other cells are some other rank's and 0 here) */
int     qhgb_get_num_agents_range(qhgb_pop *p, int32_t cell_begin, int32_t cell_end, uint64_t *out);
/* m_aiNumAgentsPerCell (core/SPopulation.cpp:236-247, updated by updateNumAgentsPerCell at the end of every step, :1250-1290) is
 * a host array that is simply current whenever the host looks at it.  The same here: `host` (cell_end - cell_begin ulongs,
 * best page-locked, qhgb_host_alloc) is refreshed by every step, event and upload under the synchronisation the call does
 * anyway -- no second round trip for the counts.  A run of queued steps (qhgb_run) refreshes it once per window.  NULL ends it. */
int     qhgb_mirror_num_agents_array(qhgb_pop *p, uint64_t *host, int32_t cell_begin, int32_t cell_end);
/* MoveStats' per-cell arrays, the datasets "Hops", "Dist", "Time" it writes into the species' action group
 * (actions/MoveStats.cpp:363-384 writeAdditionalDataQDF; m_aiHops int, m_adDist / m_adTime double, -1 = never reached) */
int     qhgb_get_move_stats(qhgb_pop *p, int32_t *hops, double *dist, double *time);
/* OccTracker::calcBitMap (core/OccTracker.cpp:95-106, called per tracked cell by updateCounts :36-45 after every step): is any agent
 * of this population in cell cells[i]?  out[i] = 1 / 0.  n bytes cross the bus instead of the whole count array; a shard answers
 * for its own cells (the host ORs the ranks' answers). */
int     qhgb_get_occupied(qhgb_pop *p, int32_t n, const int32_t *cells, uint8_t *out);
int     qhgb_get_step_stats(qhgb_pop *p, qhgb_step_stats *out);
/* parity probes for the deterministic sub-steps: the arrays the reference keeps in
 * m_adEnvWeights (n_cells*(max_neigh+1) doubles, actions/SingleEvaluator.cpp:174-243), LinearBirth::m_adB /
 * LinearDeath::m_adD (actions/LinearBirth.cpp:97-112, actions/LinearDeath.cpp:101-119), and the ATanDeath
 * probability of actions/ATanDeath.cpp:75 evaluated on the device for given float ages. */
int  qhgb_get_env_weights(qhgb_pop *p, double *out);
int  qhgb_get_birth_death_probs(qhgb_pop *p, double *b, double *d);
int  qhgb_atan_death_prob(qhgb_pop *p, int n, const float *age, double *out);
/* the carrying capacities NPPCapacity keeps in m_adCapacities (actions/NPPCapacity.cpp:138-217), one double per cell */
int  qhgb_get_capacities(qhgb_pop *p, double *out);

/* ---- dump / restore (core/SPopulation.cpp:2024-2570: dumpSpecies..., restoreSpecies...; app/Simulator.cpp dump events) ----
 * qhgb_dump_state writes the whole dynamic state of the population between two steps into one file (agents, genomes,
 * id base, the step counter that is the state of the counter-based random streams, weights/capacities and the observers'
 * flags); qhgb_restore_state loads it into a population that was created and configured like the dumped one (cells,
 * environment arrays as of the dump, attributes, priorities, navigation) but holds no agents, and replaces
 * qhgb_add_agents + qhgb_pre_loop.  The continued run is bit-identical to the uninterrupted one.  The file format is
 * this library's own (HDF5, which the reference's QDF dumps use, is outside the path).
 * `path` of qhgb_restore_state may name several files separated by '\n' -- the dumps of all ranks of a sharded run; a sharded
 * population then keeps the agents (and genome rows) of its own cell range from every file.  That is how a run is re-split
 * over new cell ranges, or another number of GPUs, when its load has shifted (after environment events):
 * every rank dumps, the host computes new ranges, new populations restore (qhg4_b200/sharding.py::rebalance). */
int  qhgb_dump_state(qhgb_pop *p, const char *path);
int  qhgb_restore_state(qhgb_pop *p, const char *path);

/* ---- several GPUs: the grid sharded by contiguous cell ranges --------------------------------------
 * One process per GPU, each with its own qhgb_pop holding the agents of its cell range [cell_begin[rank],
 * cell_begin[rank+1]); environment arrays are replicated.  Per step the ranks exchange (NCCL over NVLink) the
 * arrivals per cell, the births per rank (newborn ids are global ranks) and the packed records of the agents that
 * crossed a range boundary.  Results are identical to the single-GPU run.  The reference has no counterpart (its only
 * message-passing code is the tiling experiment tools_ico/MPIMulti.cpp:301-370); the simulator's note at
 * app/Simulator.cpp:89-101 names what a distributed run needs: the global max id and the births of the preceding nodes.
 *   qhgb_comm_get_unique_id  rank 0 creates the 128-byte NCCL id; the host distributes it (any transport)
 *   qhgb_comm_init           before qhgb_add_agents; afterwards qhgb_add_agents keeps only the agents of the own range
 *   qhgb_comm_get_traffic    agents sent / received in the last step */
int  qhgb_comm_get_unique_id(void *out, int nbytes);
int  qhgb_comm_init(qhgb_pop *p, int rank, int nranks, const void *unique_id, const int32_t *cell_begin);
int  qhgb_comm_get_traffic(qhgb_pop *p, int64_t *sent, int64_t *received);
/* Exchange over peer memory instead of NCCL calls (same results): every rank exports two device allocations (its
 * remote arrival counters and its receive buffer) as CUDA IPC handles, the host gathers the 128 bytes of every rank
 * (any transport) and hands the table to each rank.  From then on a step has no host round trip: arrival counts go
 * to the owning GPU by remote atomics, migrant records by direct NVLink stores from the scatter kernel, ordered by
 * two device-side cross-GPU barriers.
 *   qhgb_comm_p2p_handle   after qhgb_comm_init: writes 128 bytes (two cudaIpcMemHandle_t)
 *   qhgb_comm_p2p_connect  all_handles = nranks x 128 bytes in rank order; NULL switches back to the NCCL exchange
 *                          (every rank must use the same exchange) */
int  qhgb_comm_p2p_handle(qhgb_pop *p, void *out, int nbytes);
int  qhgb_comm_p2p_connect(qhgb_pop *p, const void *all_handles);

/* ---- page-locked host arrays ------------------------------------------------------------------------
 * The reference hands its per-cell arrays to the host as plain new[] memory (core/SPopulation.cpp:236-247).  An array
 * the host reads every step (qhgb_get_num_agents_array) is copied by one DMA when it is page-locked; these two calls
 * allocate and free such memory.  Every qhgb_get_* call also accepts ordinary pageable memory. */
void *qhgb_host_alloc(size_t bytes);
int   qhgb_host_free(void *ptr);

/* ---- measurement hooks ------------------------------------------------------------------------------
 * number of kernels launched by this population since creation, and the CUDA stream they run on */
/* which kernels the pipeline runs (steps, events, uploads) took so far: the fast path (one warp per batch of cells), the generic path
 * (one thread per agent: rare actions, cells beyond the fast path's limits on a single GPU), and -- sharded runs -- steps redone
 * with the recovery kernels because some rank met a cell beyond the default limits (1024 agents, 128 births, 384 ranked fertile
 * females per cell; the recovery kernels take 8192 / 2048 / 4096) */
int  qhgb_get_path_counts(qhgb_pop *p, int64_t *fast, int64_t *generic, int64_t *recovery);
int64_t qhgb_get_launch_count(qhgb_pop *p);
void   *qhgb_get_stream(qhgb_pop *p);
/* device time (ms, CUDA events on the population's stream) accumulated per kernel name since the last reset;
 * names/ms/calls hold up to cap entries; returns the number of kernels known */
int  qhgb_get_kernel_times(qhgb_pop *p, int cap, const char **names, double *ms, int64_t *calls);
int  qhgb_reset_kernel_times(qhgb_pop *p, int enable);
/* CUDA events on the population's stream: record into slot 0..7, elapsed ms between two recorded slots (-1 on error) */
int    qhgb_event_record(qhgb_pop *p, int slot);
double qhgb_event_elapsed_ms(qhgb_pop *p, int slot_a, int slot_b);

#ifdef __cplusplus
}
#endif
#endif
