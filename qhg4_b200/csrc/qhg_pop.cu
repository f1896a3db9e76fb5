// qhg_pop.cu -- host side of the C ABI declared in include/qhg_b200.h.
//
// A qhgb_pop owns the device-resident population (structure of arrays, binned by cell, double buffered),
// the per-cell arrays and one CUDA stream.  The host object mirrors what SPopulation<T> + Prioritizer<T> +
// the Action<T> objects hold on the host in the reference (core/SPopulation.h:86-344,
// core/Prioritizer.h:19-78): named actions with priorities and enable flags, named attributes, per-step
// totals.  No CPU fallback exists: every entry point that computes launches kernels from qhg_kernels.cuh.
#include "../../include/qhg_b200.h"
#include "qhg_cells.cuh"
#include "qhg_decide.cuh"
#include "qhg_genes.cuh"

#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

using namespace qhg;

namespace {

thread_local std::string g_err;

int fail(const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return -1;
}

// no exception crosses the C boundary (include/qhg_b200.h): entry points that size host containers from their arguments or
// from a file run under this guard
template <class F>
auto guarded(F f) -> decltype(f()) {
    try {
        return f();
    } catch (const std::exception &e) {
        return (decltype(f()))fail("%s", e.what());
    } catch (...) {
        return (decltype(f()))fail("unknown C++ exception");
    }
}

#define CK(call)                                                                                          \
    do {                                                                                                  \
        cudaError_t e_ = (call);                                                                          \
        if (e_ != cudaSuccess) return fail("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

// NCCL is bound at run time (dlopen) so that the library also loads on hosts without it; the communicator is only
// needed when a population is sharded over several GPUs (qhgb_comm_init).
struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
};
NcclApi g_nccl;

int loadNccl() {
    if (g_nccl.handle) return 0;
    const char *names[] = {getenv("QHG_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    void *h = nullptr;
    for (const char *n : names) {
        if (!n || !*n) continue;
        h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) return fail("NCCL library not found (set QHG_NCCL_LIB): %s", dlerror());
#define QHG_SYM(field, name)                                                                 \
    *(void **)(&g_nccl.field) = dlsym(h, name);                                              \
    if (!g_nccl.field) return fail("NCCL symbol %s missing", name)
    QHG_SYM(GetUniqueId, "ncclGetUniqueId");
    QHG_SYM(CommInitRank, "ncclCommInitRank");
    QHG_SYM(CommDestroy, "ncclCommDestroy");
    QHG_SYM(GetErrorString, "ncclGetErrorString");
    QHG_SYM(AllReduce, "ncclAllReduce");
    QHG_SYM(AllGather, "ncclAllGather");
    QHG_SYM(Send, "ncclSend");
    QHG_SYM(Recv, "ncclRecv");
    QHG_SYM(GroupStart, "ncclGroupStart");
    QHG_SYM(GroupEnd, "ncclGroupEnd");
#undef QHG_SYM
    g_nccl.handle = h;
    return 0;
}

#define NK(call)                                                                                           \
    do {                                                                                                   \
        ncclResult_t r_ = (call);                                                                          \
        if (r_ != ncclSuccess) return fail("%s failed: %s (%s:%d)", #call, g_nccl.GetErrorString(r_), __FILE__, __LINE__); \
    } while (0)

enum ActKind { A_GETOLD, A_ATANDEATH, A_OLDAGEDEATH, A_WEIGHTEDMOVE, A_SINGLEEVAL, A_FERTILITY, A_RANDOMPAIR, A_VERHULST,
               A_VERHULSTVARK, A_MULTIEVAL, A_NPPCAP, A_GENETICS, A_NAVIGATE, A_RANDOMMOVE, A_CONFINEDMOVE, A_WEIGHTEDMOVERAND, A_SIGDEATH,
               A_CONDWEIGHTEDMOVE, A_RANDPERMPAIR, A_MOVESTATS };

// one SingleEvaluator inside a MultiEvaluator (actions/SingleEvaluator.cpp:138-167)
struct SubEval {
    std::string input;       // environment array it evaluates, "" = the capacities array of NPPCapacity
    std::string weightName;  // attribute holding its combination weight
    bool usePoly = false;
    std::string polyName;    // attribute holding its poly-line
    int trigger = 0;         // event id that makes it recompute
    bool first = true, needUpdate = false;
    bool cumulate = true;    // the bCumulate argument of its constructor (actions/SingleEvaluator.cpp:216-243)
};

struct HostAction {
    std::string name;
    ActKind kind;
    int prio = -1;  // no <prio> entry: the action exists but is never run (core/Prioritizer.cpp:19-30)
    bool enabled = true;
};

struct KernelTime {
    std::string name;
    double ms = 0;
    int64_t calls = 0;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pending;
};

template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    cudaError_t alloc(size_t count) {
        if (p) cudaFree(p);
        p = nullptr;
        n = count;
        if (count == 0) return cudaSuccess;
        cudaError_t e = cudaMalloc(&p, count * sizeof(T));
        if (e != cudaSuccess) return e;
        // per-cell counters of cells another rank owns are never touched again.  The zeroing runs on the default stream, which
        // the populations' non-blocking streams do not wait for: it must have finished before anybody uses the buffer
        e = cudaMemset(p, 0, count * sizeof(T));
        if (e != cudaSuccess) return e;
        return cudaStreamSynchronize(0);
    }
    // room for `count` elements whose old contents do not matter: the allocation is kept when it is large enough (tables that
    // are rebuilt on every environment event; cudaFree / cudaMalloc inside a run cost milliseconds)
    cudaError_t reserve(size_t count) {
        if (p && cap >= count) { n = count; return cudaSuccess; }
        cudaError_t e = alloc(count);
        if (e == cudaSuccess) cap = count;
        return e;
    }
    size_t cap = 0;
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
        cap = 0;
    }
};

}  // namespace

struct qhgb_pop {
    std::string popClass;
    int device = 0;
    int nCells = 0, maxNeigh = 6;
    int numSMs = 148;
    cudaStream_t stream = nullptr;

    std::vector<HostAction> actions;
    std::map<std::string, double> attr;
    std::map<std::string, std::string> attrStr;
    PolyLineDev poly{};
    bool havePoly = false;
    uint32_t state16[16] = {0};
    RngKey key{0, 0};

    // grid / env
    DevBuf<int> nbr, gid, count[2], cellStart[2], stay, arrive, cursor, birthCount, birthBase, nFert;
    DevBuf<unsigned long long> count64;  // per-cell counts widened for qhgb_get_num_agents_array
    uint64_t *mirror = nullptr;          // host array kept current after every step (qhgb_mirror_num_agents_array), cells [mirrorLo, mirrorHi)
    int mirrorLo = 0, mirrorHi = 0;
    cudaStream_t copyStream = nullptr;   // the mirror's copy runs beside the scatter pass: the counts are final after the scan
    cudaEvent_t evScan = nullptr, evCopy = nullptr;
    bool mirrorQueued = false;
    DevBuf<int> occCells;                // qhgb_get_occupied: the tracked cells and their answer
    DevBuf<uint8_t> occOut;
    DevBuf<int> moveBase;  // fast path: first arrival slot of the movers of (cell, direction), MOVE_STRIDE ints per cell
    DevBuf<uint8_t> nNbr, ice;
    DevBuf<double> alt, W, B, D;
    DevBuf<unsigned long long> TB, TD;  // B, D as integer thresholds for the fast path (k_cell_init)
    DevBuf<int2> tileSums;
    std::map<std::string, DevBuf<double>> envExtra;
    bool haveCells = false, haveAlt = false, haveIce = false;
    std::vector<int32_t> hGid;

    // agents
    int64_t capacity = 0;
    DevBuf<int64_t> id[2];
    DevBuf<float> birth[2], lastBirth[2], age[2];
    DevBuf<int> cell[2], mate, prank, ranked, dest, rank;
    DevBuf<uint8_t> flags[2], oflags, dec;
    DevBuf<uint32_t> pkey;
    int cur = 0;
    DevBuf<DevStats> dstats;
    DevStats *hstats = nullptr;  // pinned
    // m_fAge: either the device array holds it (ageValid) or, when the action set refreshes the age from the birth
    // time before anything reads it, it is implied: age = lastAgeTime - birth (see programNeedsStoredAge)
    bool ageValid = true;
    float lastAgeTime = 0;

    // step state
    bool preLooped = false, inStep = false, pairingValid = false, needPair = false, doVerhulst = false;
    bool forceBig = false;  // QHG_FORCE_BIG=1: sharded runs use the recovery kernels in every step (tests)
    int64_t bigSteps = 0;   // steps a sharded run had to redo with the recovery kernels
    bool forceGeneric = false;
    bool cellValid = true;  // does cell[cur] hold the per-agent cell index? (the fast path leaves it implied by cellStart)
    int64_t genericSteps = 0, tiledSteps = 0;
    bool evalFirst = true, evalNeedUpdate = false;
    bool evaluatorObserves = false;  // does the population class addObserver() its evaluator?
    // tut_EnvironCapAltPop: NPPCapacity + MultiEvaluator[NPP+Alt] + VerhulstVarK (populations/tut_EnvironCapAltPop.cpp:27-72)
    std::vector<SubEval> subs;
    bool multiFirst = true, nppNeedUpdate = true, multiObserves = false;
    DevBuf<double> cap, Wtmp;
    int multiMode = MM_ADD_SIMPLE;   // combine mode of the class's MultiEvaluator (actions/MultiEvaluator.h:19-26)
    DevBuf<uint8_t> multiAllowed;    // MultiEvaluator::m_acAllowed (the BLOCK modes)
    std::map<std::string, PolyLineDev> polys;
    // Genetics<.., BitGeneUtils>: genome pool (rows are not moved by the re-binning), free stack, birth list
    bool genetic = false;
    GeneParams gp{};
    DevBuf<int> gslot[2], nbabies[2], gfree;
    DevBuf<unsigned long long> gpool;
    DevBuf<BirthEntry> births;
    DevBuf<int> father;      // fast path with Genetics: position of the mate of every mother-to-be (k_cell_decide<false, true>)
    bool genFast = false;    // populations with Genetics take the fast path (QHG_GEN_FAST=0 switches it off)
    bool navFast = false;    // programs that end with Navigate take the fast path (QHG_NAV_FAST=0 switches it off)
    bool segDecide = true;   // pass 1 by batches of cells (qhg_decide.cuh); QHG_DECIDE=cell: one warp per cell (qhg_cells.cuh)
    bool drownsOnGeo = false; // does the class override updateEvent to kill the agents of flooded / iced cells?
    DevBuf<JumpEntry> jumps; // fast path with Navigate: the agents that jump this step
    DevBuf<int> jumpCount;
    DevBuf<GenomeCtl> gctl;
    int64_t poolRows = 0;
    // Navigate: the Navigation group as the host handed it over, and the jump tables built from it
    std::vector<int> hPortCell, hPortPtr, hDestCell, hBridges;
    std::vector<double> hDist;
    bool navNeedUpdate = true, navReady = false;
    int nCurBridges = 0;
    DevBuf<int> navRow, navPtr, navDest;
    DevBuf<double> navCum;
    DevBuf<int2> navBridges;
    std::vector<double> hAlt;  // host copy of the altitude (bridges need both ends above sea level)
    bool hAltStale = false;    // the device array was interpolated since the copy was taken
    std::map<std::string, DevBuf<double>> envDelta;  // per-step difference arrays of the interpolated targets (AutoInterpolator::m_mDiff)
    // ConfinedMove: the cells inside the region (ConfinedMove::m_bAllowed, actions/ConfinedMove.cpp:44-78), built at preLoop
    DevBuf<uint8_t> allowed;
    bool confReady = false;
    // MoveStats (actions/MoveStats.cpp): main arrays, the reference's Temp arrays, the per-step atomics (qhg_kernels.cuh)
    DevBuf<int> msHops, msHopsT, msStepHops;
    DevBuf<double> msDist, msTime, msDistT, msTimeT, msLon, msLat;
    DevBuf<unsigned long long> msKey, msStepDist;
    DevBuf<uint8_t> msChanged;
    DevBuf<MoveStatsDev> msDev;
    MoveStatsDev msHost{};
    bool msReady = false;
    int condMode = 0;  // SimpleCondition mode of a CondWeightedMove (the probe classes tut_EnvironAltCond<m>Pop carry it in their name)
    std::vector<double> hLon, hLat;  // host copies of Longitude / Latitude for it
    bool selfMate = false;           // tut_ParthenoPop: every female counts as mated, newborns are female
    float curTime = -1;
    std::vector<unsigned> levels;
    int64_t nAgents = 0, maxID = 0, stepsDone = 0;
    int64_t lastBirths = 0, lastDeaths = 0, lastMoves = 0, nextID = 0;
    int64_t launches = 0;

    // multi-GPU sharding (qhgb_comm_init): contiguous cell ranges, one per rank
    bool sharded = false;
    int shRank = 0, shRanks = 1;
    std::vector<int> cellBegin;
    std::vector<int> hNbr;  // host copy of the neighbour table (halo construction)
    ncclComm_t comm = nullptr;
    DevBuf<int> dCellBegin, dInfo, dAllInfo, dSendOff, dSendCursor;
    DevBuf<int> dHalo, dHaloBuf;  // cells with a neighbour on another rank (same list on all ranks) and their exchanged arrival counts
    int nHalo = 0;
    int cLo() const { return sharded ? cellBegin[shRank] : 0; }
    int cHi() const { return sharded ? cellBegin[shRank + 1] : nCells; }
    DevBuf<Migrant> sendBuf, recvBuf;
    DevBuf<unsigned long long> sendGenomes, recvGenomes;  // NCCL exchange of populations with Genetics: the migrants' genome rows
    int *hAllInfo = nullptr;  // pinned: nranks * (nranks + 1) ints
    int64_t lastSent = 0, lastReceived = 0;
    int64_t agentSteps = 0, totSent = 0, totRecv = 0;  // host mirrors of the device-side sums (DevStats)
    // exchange over peer memory (qhgb_comm_p2p_handle / qhgb_comm_p2p_connect)
    bool p2p = false;
    int *arriveRemote = nullptr;     // [2][nCells], written by the peers
    XchgBlock *xchg = nullptr;       // header + receive buffer, written by the peers
    int recvCap = 0;
    unsigned xStep = 0;              // exchanges done so far (the same on every rank): parity and barrier stamp
    DevBuf<int> remoteBase;
    DevBuf<PeerTable> dPeers;
    std::vector<void *> ipcOpened;

    bool timing = false;
    cudaEvent_t userEv[8] = {nullptr};
    std::vector<KernelTime> ktimes;

    AgentArrays arrays(int b) {
        return AgentArrays{id[b].p, birth[b].p, lastBirth[b].p, cell[b].p, flags[b].p, age[b].p, genetic ? gslot[b].p : nullptr,
                           genetic ? nbabies[b].p : nullptr};
    }
    HostAction *findKind(ActKind k) {
        for (auto &a : actions) if (a.kind == k) return &a;
        return nullptr;
    }
    bool active(ActKind k) {
        HostAction *a = findKind(k);
        return a && a->prio >= 0 && a->enabled;
    }
    HostAction *find(const std::string &n) {
        for (auto &a : actions) if (a.name == n) return &a;
        return nullptr;
    }
    double A(const char *n, double def = 0) const {
        auto it = attr.find(n);
        return it == attr.end() ? def : it->second;
    }
    int gridFor(int64_t n, int block = 256, int perSM = 8) const {
        int64_t need = (n + block - 1) / block;
        int64_t cap = (int64_t)numSMs * perSM;
        return (int)std::max<int64_t>(1, std::min(need, cap));
    }
    KernelTime &kt(const char *name) {
        for (auto &k : ktimes) if (k.name == name) return k;
        ktimes.push_back(KernelTime{name});
        return ktimes.back();
    }
};

#define LAUNCH(p, name, kern, grid, block, ...)                                \
    do {                                                                       \
        cudaEvent_t e0_ = nullptr, e1_ = nullptr;                              \
        if ((p)->timing) {                                                     \
            cudaEventCreate(&e0_);                                             \
            cudaEventCreate(&e1_);                                             \
            cudaEventRecord(e0_, (p)->stream);                                 \
        }                                                                      \
        kern<<<(grid), (block), 0, (p)->stream>>>(__VA_ARGS__);                \
        (p)->launches++;                                                       \
        if ((p)->timing) {                                                     \
            cudaEventRecord(e1_, (p)->stream);                                 \
            (p)->kt(name).pending.push_back({e0_, e1_});                       \
        }                                                                      \
    } while (0)

#define LAUNCH_SMEM(p, name, kern, grid, block, smem, ...)                     \
    do {                                                                       \
        cudaEvent_t e0_ = nullptr, e1_ = nullptr;                              \
        if ((p)->timing) {                                                     \
            cudaEventCreate(&e0_);                                             \
            cudaEventCreate(&e1_);                                             \
            cudaEventRecord(e0_, (p)->stream);                                 \
        }                                                                      \
        kern<<<(grid), (block), (smem), (p)->stream>>>(__VA_ARGS__);           \
        (p)->launches++;                                                       \
        if ((p)->timing) {                                                     \
            cudaEventRecord(e1_, (p)->stream);                                 \
            (p)->kt(name).pending.push_back({e0_, e1_});                       \
        }                                                                      \
    } while (0)

// any stream-ordered statement (NCCL calls, copies) under the same optional event timing as the kernels
#define TIMED(p, name, ...)                                                    \
    do {                                                                       \
        cudaEvent_t e0_ = nullptr, e1_ = nullptr;                              \
        if ((p)->timing) {                                                     \
            cudaEventCreate(&e0_);                                             \
            cudaEventCreate(&e1_);                                             \
            cudaEventRecord(e0_, (p)->stream);                                 \
        }                                                                      \
        __VA_ARGS__;                                                           \
        if ((p)->timing) {                                                     \
            cudaEventRecord(e1_, (p)->stream);                                 \
            (p)->kt(name).pending.push_back({e0_, e1_});                       \
        }                                                                      \
    } while (0)

namespace {

int allocAgents(qhgb_pop *p, int64_t cap) {
    // grow (or create) the agent buffers, keeping the live prefix of the current buffer
    if (cap > (int64_t)2000000000) return fail("capacity %lld exceeds 32-bit agent indices", (long long)cap);
    qhgb_pop &q = *p;
    for (int b = 0; b < 2; b++) {
        int64_t keep = (b == q.cur) ? q.nAgents : 0;
        auto regrow = [&](auto &buf) -> cudaError_t {
            using T = std::remove_pointer_t<decltype(buf.p)>;
            T *np = nullptr;
            cudaError_t e = cudaMalloc(&np, (cap + AGENT_SLACK) * sizeof(T));  // bulk copies read whole 16-agent groups
            if (e != cudaSuccess) return e;
            if (keep > 0 && buf.p) e = cudaMemcpyAsync(np, buf.p, keep * sizeof(T), cudaMemcpyDeviceToDevice, q.stream);
            cudaStreamSynchronize(q.stream);
            if (buf.p) cudaFree(buf.p);
            buf.p = np;
            buf.n = cap;
            return e;
        };
        CK(regrow(q.id[b]));
        CK(regrow(q.birth[b]));
        CK(regrow(q.lastBirth[b]));
        CK(regrow(q.age[b]));
        CK(regrow(q.cell[b]));
        CK(regrow(q.flags[b]));
        if (q.genetic) {
            CK(regrow(q.gslot[b]));
            CK(regrow(q.nbabies[b]));
        }
    }
    if (q.genetic) {  // the pool keeps its rows; the free stack its entries
        const size_t row = 2 * (size_t)q.gp.nBlocks;
        unsigned long long *np = nullptr;
        CK(cudaMalloc(&np, (size_t)cap * row * sizeof(unsigned long long)));
        if (q.gpool.p && q.poolRows > 0) CK(cudaMemcpyAsync(np, q.gpool.p, (size_t)q.poolRows * row * sizeof(unsigned long long), cudaMemcpyDeviceToDevice, q.stream));
        int *nf = nullptr;
        CK(cudaMalloc(&nf, (size_t)cap * sizeof(int)));
        if (q.gfree.p && q.poolRows > 0) CK(cudaMemcpyAsync(nf, q.gfree.p, (size_t)q.poolRows * sizeof(int), cudaMemcpyDeviceToDevice, q.stream));
        CK(cudaStreamSynchronize(q.stream));
        if (q.gpool.p) cudaFree(q.gpool.p);
        if (q.gfree.p) cudaFree(q.gfree.p);
        q.gpool.p = np; q.gpool.n = (size_t)cap * row;
        q.gfree.p = nf; q.gfree.n = cap;
        q.poolRows = cap;
        CK(q.births.alloc(cap / 2 + 1024));
        if (q.genFast) CK(q.father.alloc(cap));
    }
    CK(q.mate.alloc(cap));
    CK(q.prank.alloc(cap));
    CK(q.ranked.alloc(cap));
    CK(q.dest.alloc(cap));
    CK(q.rank.alloc(cap));
    CK(q.oflags.alloc(cap));
    CK(q.dec.alloc(cap + AGENT_SLACK));
    CK(q.pkey.alloc(cap));
    q.capacity = cap;
    q.pairingValid = false;
    return 0;
}

int ensureCapacity(qhgb_pop *p, int64_t need) {
    if (need <= p->capacity) return 0;
    int64_t cap = std::max<int64_t>(need + need / 4, 1024);
    return allocAgents(p, cap);
}

int pushStats(qhgb_pop *p) {
    DevStats s{};
    s.nAgents = (int)p->nAgents;
    s.nextID = p->nextID;
    s.step = (unsigned)p->stepsDone;
    s.agentSteps = p->agentSteps; s.totSent = p->totSent; s.totRecv = p->totRecv;
    CK(cudaMemcpyAsync(p->dstats.p, &s, sizeof(s), cudaMemcpyHostToDevice, p->stream));
    CK(cudaStreamSynchronize(p->stream));
    return 0;
}

void enqueueMirror(qhgb_pop *p, int buf) {
    LAUNCH(p, "k_counts_u64", k_counts_u64, p->gridFor(p->mirrorHi - p->mirrorLo), 256, p->mirrorLo, p->mirrorHi, p->cLo(), p->cHi(), p->count[buf].p,
           p->count64.p);
    cudaMemcpyAsync(p->mirror, p->count64.p + p->mirrorLo, sizeof(uint64_t) * (size_t)(p->mirrorHi - p->mirrorLo), cudaMemcpyDeviceToHost, p->stream);
}

// inside a step: the next counts are final once the scan has run (every arrival, birth and death is decided in pass 1), so the
// host's copy is made on a second stream WHILE pass 2 moves the agents; the step's stream waits for it before the host does
void mirrorAfterScan(qhgb_pop *p, int buf) {
    cudaEventRecord(p->evScan, p->stream);
    cudaStreamWaitEvent(p->copyStream, p->evScan, 0);
    k_counts_u64<<<p->gridFor(p->mirrorHi - p->mirrorLo), 256, 0, p->copyStream>>>(p->mirrorLo, p->mirrorHi, p->cLo(), p->cHi(), p->count[buf].p, p->count64.p);
    p->launches++;
    cudaMemcpyAsync(p->mirror, p->count64.p + p->mirrorLo, sizeof(uint64_t) * (size_t)(p->mirrorHi - p->mirrorLo), cudaMemcpyDeviceToHost, p->copyStream);
    cudaEventRecord(p->evCopy, p->copyStream);
    p->mirrorQueued = true;
}

int pullStats(qhgb_pop *p) {
    CK(cudaMemcpyAsync(p->hstats, p->dstats.p, sizeof(DevStats), cudaMemcpyDeviceToHost, p->stream));
    CK(cudaStreamSynchronize(p->stream));
    return 0;
}

// derive the fused action program from the host-side action table: priority ascending, same priority in
// name order (core/SPopulation.cpp:249-257 iterates a std::map<string,int>), only the requested levels
ActParams buildProgram(qhgb_pop *p, const std::vector<unsigned> *levels, float t) {
    ActParams P{};
    std::vector<const HostAction *> v;
    for (auto &a : p->actions) {
        if (a.prio < 0 || !a.enabled) continue;
        if (levels && std::find(levels->begin(), levels->end(), (unsigned)a.prio) == levels->end()) continue;
        v.push_back(&a);
    }
    std::stable_sort(v.begin(), v.end(), [](const HostAction *x, const HostAction *y) {
        if (x->prio != y->prio) return x->prio < y->prio;
        return x->name < y->name;
    });
    for (const HostAction *a : v) {
        uint8_t op = 0;
        switch (a->kind) {
        case A_GETOLD: op = OP_GETOLD; break;
        case A_ATANDEATH: op = OP_ATANDEATH; break;
        case A_OLDAGEDEATH: op = OP_OLDAGEDEATH; break;
        case A_WEIGHTEDMOVE: op = OP_WEIGHTEDMOVE; break;
        case A_RANDOMMOVE: op = OP_RANDOMMOVE; break;
        case A_WEIGHTEDMOVERAND: op = OP_WEIGHTEDMOVERAND; break;
        case A_CONDWEIGHTEDMOVE: op = OP_CONDWEIGHTEDMOVE; break;
        case A_SIGDEATH: op = OP_SIGDEATH; break;
        case A_FERTILITY: op = OP_FERTILITY; break;
        case A_VERHULST: op = OP_VERHULST; break;
        case A_VERHULSTVARK: op = OP_VERHULST; break;
        case A_NAVIGATE: op = OP_NAVIGATE; break;  // the same two executes with a per-cell K (actions/VerhulstVarK.cpp:96-112)
        default: break;  // evaluators and pairing have no per-agent execute()
        }
        if (op && P.nOps < MAX_OPS) { P.prog |= (unsigned long long)op << (4 * P.nOps); P.nOps++; }
    }
    P.t = t;
    P.storeAge = 1;
    P.key = p->key;
    for (int r = 0; r < 10; r++) { P.rk.a[r] = p->key.k0 + (uint32_t)r * 0x9E3779B9u; P.rk.b[r] = p->key.k1 + (uint32_t)r * 0xBB67AE85u; }
    // ATanDeath::preLoop, actions/ATanDeath.cpp:49-59 (EPS = 0.001, actions/ATanDeath.h:14)
    P.atanMaxAge = p->A("ATanDeath_max_age");
    P.atanSlope = p->A("ATanDeath_slope");
    double range = p->A("ATanDeath_range");
    P.atanScale = (M_PI / 2 - 0.001) / atan(P.atanSlope * range);
    // outside [Xlo, Xhi] the probability is < 0 resp. > 1 whatever the draw: skip the atan there
    P.atanXlo = -INFINITY;
    P.atanXhi = INFINITY;
    if (P.atanScale > 1.0 + 1e-9 && std::isfinite(P.atanScale)) {
        double th = tan(M_PI / (2 * P.atanScale));
        P.atanXhi = th * (1 + 1e-6) + 1e-9;
        P.atanXlo = -P.atanXhi;
    }
    // the same window on the float age, widened by a safety margin, for the fast path's cheap pre-test
    P.atanAgeLo = -INFINITY;
    P.atanAgeHi = INFINITY;
    if (std::isfinite(P.atanXhi) && P.atanSlope > 0) {
        double lo = P.atanMaxAge + P.atanXlo / P.atanSlope, hi = P.atanMaxAge + P.atanXhi / P.atanSlope;
        double m = 1e-3 * (fabs(lo) + fabs(hi) + 1.0);
        P.atanAgeLo = nextafterf((float)(lo - m), -INFINITY);
        P.atanAgeHi = nextafterf((float)(hi + m), INFINITY);
    }
    P.oadMaxAge = p->A("OAD_max_age");
    double unc = p->A("OAD_uncertainty");
    P.oadLo = 1 - unc * P.oadMaxAge;
    P.oadHi = 1 + unc * P.oadMaxAge;
    P.moveProb = p->findKind(A_RANDOMMOVE) ? p->A("RandomMove_prob") : p->A("WeightedMove_prob");  // a population has one move action
    if (p->active(A_CONDWEIGHTEDMOVE)) P.moveProb = p->A("CondWeightedMove_prob");  // (the probe classes carry two: the <prio> entries choose)
    P.condMode = p->condMode;
    P.tMove = (P.moveProb > 0) ? (unsigned long long)ceil(P.moveProb * 4294967296.0) : 0ull;  // exact: a power-of-two scaling
    P.fertMinAge = (float)p->A("Fertility_min_age");
    P.fertMaxAge = (float)p->A("Fertility_max_age");
    P.fertInterbirth = (float)p->A("Fertility_interbirth");
    P.moveProbRand = p->A("WeightedMoveRand_prob");
    P.sigMaxAge = p->A("SigDeath_max_age");
    P.sigSlope = p->A("SigDeath_slope");
    P.sigScale = 1 + exp(-p->A("SigDeath_range"));  // SigDeath::preLoop, actions/SigDeath.cpp:49-60
    P.selfMate = p->selfMate ? 1 : 0;
    P.confine = (p->active(A_CONFINEDMOVE) && p->confReady) ? 1 : 0;  // its finalize() runs in every finalizeStep, whatever the levels
    return P;
}

// does the program refresh m_fAge from the birth time before anything reads it?  Then the age never has to be
// stored or moved: it is (time of the last refresh - birth time), also for the records handed back to the host.
bool programNeedsStoredAge(const ActParams &P) {
    for (int k = 0; k < P.nOps; k++) {
        switch (prog_op(P, k)) {
        case OP_GETOLD: case OP_ATANDEATH: case OP_OLDAGEDEATH: case OP_SIGDEATH: return false;
        case OP_FERTILITY: return true;
        default: break;
        }
    }
    return true;  // nothing touches the age: keep what is stored
}

int materializeAges(qhgb_pop *p) {
    if (p->ageValid) return 0;
    if (p->nAgents > 0) {
        LAUNCH(p, "k_fill_age", k_fill_age, p->gridFor(p->nAgents), 256, p->dstats.p, p->birth[p->cur].p, p->age[p->cur].p, p->lastAgeTime);
        CK(cudaGetLastError());
    }
    p->ageValid = true;
    return 0;
}

// NPPCapacity::recalculate on the device (actions/NPPCapacity.cpp:138-217)
int recalcCapacities(qhgb_pop *p) {
    qhgb_pop &q = *p;
    if (!q.nppNeedUpdate) return 0;
    const char *need[] = {"AnnualMeanTemp", "AnnualRainfall", "Water", "BaseNPP", "Longitude", "Latitude", "Coastal"};
    for (const char *n : need) {
        if (!q.envExtra.count(n)) {  // an array that was never set is all zero (Geography::init memsets them)
            CK(q.envExtra[n].alloc(q.nCells));
            CK(cudaMemsetAsync(q.envExtra[n].p, 0, sizeof(double) * q.nCells, q.stream));
        }
    }
    if (!q.haveAlt) return fail("[NPPCapacity] no geography (Altitude)");
    NppParams Q{q.A("NPPCap_water_factor"), q.A("NPPCap_coastal_factor"), q.A("NPPCap_coastal_min_latitude"), q.A("NPPCap_coastal_max_latitude"),
                q.A("NPPCap_NPP_min"), q.A("NPPCap_NPP_max"), q.A("NPPCap_K_max"), q.A("NPPCap_K_min"), q.A("NPPCap_efficiency", 1.0)};
    LAUNCH(p, "k_npp_capacity", k_npp_capacity, q.gridFor(q.nCells), 256, q.nCells, Q, q.envExtra["AnnualMeanTemp"].p, q.envExtra["AnnualRainfall"].p,
           q.envExtra["Water"].p, q.envExtra["BaseNPP"].p, q.alt.p, q.envExtra["Longitude"].p, q.envExtra["Latitude"].p, q.envExtra["Coastal"].p, q.cap.p);
    CK(cudaGetLastError());
    q.nppNeedUpdate = false;
    return 0;
}

// SingleEvaluator::initialize inside a MultiEvaluator (actions/SingleEvaluator.cpp:138-167): the evaluators share ONE scratch array
// (Wtmp = m_adSingleEvalWeights); it is rewritten only when the evaluator needs an update or has never run
int subEvalInit(qhgb_pop *p, SubEval &e) {
    qhgb_pop &q = *p;
    if (!(e.needUpdate || e.first)) return 0;
    e.first = false;
    const size_t nW = (size_t)q.nCells * WSTRIDE;
    const int g = q.gridFor(q.nCells);
    const double *in = e.input.empty() ? q.cap.p : (e.input == "Altitude" ? q.alt.p : q.envExtra[e.input].p);
    if (!in) return fail("No array with name [%s] found", e.input.c_str());
    const bool havePl = e.usePoly && q.polys.count(e.polyName);
    CK(cudaMemsetAsync(q.Wtmp.p, 0, nW * sizeof(double), q.stream));  // calcValues starts with a memset (:175)
    LAUNCH(p, "k_weights_own", k_weights_own, g, 256, q.nCells, in, q.haveIce ? q.ice.p : nullptr, havePl ? q.polys[e.polyName] : q.poly, havePl ? 1 : 0, q.Wtmp.p);
    LAUNCH(p, "k_weights_cumulate", k_weights_cumulate, g, 256, q.nCells, q.nbr.p, q.Wtmp.p, e.cumulate ? 1 : 0);
    return 0;
}

// MultiEvaluator::initialize and its six combine modes (actions/MultiEvaluator.cpp:142-182; 221-253 ADD_SIMPLE, 263-298 ADD_BLOCK,
// 308-342 MUL_SIMPLE, 350-381 MAX_SIMPLE -- the only one whose rows are not cumulated --, 391-432 MAX_BLOCK, 440-478 MIN_SIMPLE;
// findBlockings 579-598).  An evaluator that needs no update contributes the zeros of the memset before its initialize; in
// findBlockings there is no memset, so such an evaluator is judged by the scratch array as the previous user left it.
int computeMultiWeights(qhgb_pop *p) {
    qhgb_pop &q = *p;
    bool need = false;
    for (auto &e : q.subs) need |= e.needUpdate;
    if (!(need || q.multiFirst)) return 0;
    q.multiFirst = false;
    const size_t nW = (size_t)q.nCells * WSTRIDE;
    const int g = q.gridFor(q.nCells), gw = q.gridFor((int64_t)nW);
    const int mode = q.multiMode;
    const bool block = mode == MM_ADD_BLOCK || mode == MM_MAX_BLOCK;
    if (block) {
        if (q.multiAllowed.n != nW) CK(q.multiAllowed.alloc(nW));
        CK(cudaMemsetAsync(q.multiAllowed.p, 1, nW, q.stream));
        for (auto &e : q.subs) {
            if (subEvalInit(p, e) != 0) return -1;
            LAUNCH(p, "k_find_blockings", k_find_blockings, gw, 256, nW, q.Wtmp.p, q.multiAllowed.p);
        }
    }
    const double init = (mode == MM_MUL_SIMPLE) ? 1.0 : (mode == MM_MAX_SIMPLE || mode == MM_MAX_BLOCK) ? -INFINITY : (mode == MM_MIN_SIMPLE) ? INFINITY : 0.0;
    LAUNCH(p, "k_fill_f64", k_fill_f64, gw, 256, nW, init, q.W.p);
    for (auto &e : q.subs) {
        CK(cudaMemsetAsync(q.Wtmp.p, 0, nW * sizeof(double), q.stream));
        if (subEvalInit(p, e) != 0) return -1;
        LAUNCH(p, "k_multi_combine", k_multi_combine, gw, 256, nW, mode, q.Wtmp.p, q.A(e.weightName.c_str()), block ? q.multiAllowed.p : nullptr, q.W.p);
    }
    if (mode != MM_MAX_SIMPLE) LAUNCH(p, "k_rows_cumulate", k_rows_cumulate, g, 256, q.nCells, q.W.p);
    CK(cudaGetLastError());
    return 0;
}

// mutation-count table of Genetics (utils/BinomialDist.cpp:61-87): table[k] = 1 - I_p(k+1, n-k) until the tail is below eps;
// the incomplete beta function by the classic log-gamma series + continued fraction (utils/bino_tools.cpp)
double gammaLn(double xx) {
    static const double co[6] = {76.18009172947146, -86.50532032941677, 24.01409824083091, -1.231739572450155, 0.1208650973866179e-2, -0.5395239384953e-5};
    double ser = 1.000000000190015, x = xx, y = xx + 1, tmp = x + 5.5;
    tmp -= (x + 0.5) * log(tmp);
    for (int k = 0; k <= 5; k++) { ser += co[k] / y; y++; }
    return -tmp + log(2.5066282746310005 * ser / x);
}
double betaCf(double a, double b, double x) {
    const double eps = 3.0e-7; const float fmin_ = 1.0e-30f;
    double qab = a + b, qap = a + 1.0, qam = a - 1.0, c = 1.0, d = 1.0 - qab * x / qap;
    if (fabs(d) < fmin_) d = fmin_;
    d = 1.0 / d;
    double h = d;
    int m;
    for (m = 1; m <= 100; m++) {
        int m2 = 2 * m;
        double aa = m * (b - m) * x / ((qam + m2) * (a + m2));
        d = 1.0 + aa * d; if (fabs(d) < fmin_) d = fmin_;
        c = 1.0 + aa / c; if (fabs(c) < fmin_) c = fmin_;
        d = 1.0 / d; h *= d * c;
        aa = -(a + m) * (qab + m) * x / ((a + m2) * (qap + m2));
        d = 1.0 + aa * d; if (fabs(d) < fmin_) d = fmin_;
        c = 1.0 + aa / c; if (fabs(c) < fmin_) c = fmin_;
        d = 1.0 / d;
        double del = d * c;
        h *= del;
        if (fabs(del - 1.0) < eps) break;
    }
    return (m > 100) ? -1 : h;
}
double incBeta(double a, double b, double x) {
    if (x < 0.0 || x > 1.0) return -1;
    double bt = (x == 0.0 || x == 1.0) ? 0.0 : exp(gammaLn(a + b) - gammaLn(a) - gammaLn(b) + a * log(x) + b * log(1.0 - x));
    if (x < (a + 1.0) / (a + b + 2.0)) return bt * betaCf(a, b, x) / a;
    return 1.0 - bt * betaCf(b, a, 1.0 - x) / b;
}
std::vector<double> binomialTable(double prob, int n, double eps) {
    std::vector<double> v;
    int k = 1;
    double d2 = incBeta(k, n - k + 1, prob);
    while (d2 > eps && k < n) { v.push_back(d2); k++; d2 = incBeta(k, n - k + 1, prob); }
    v.push_back(d2);
    for (auto &x : v) x = 1 - x;
    return v;
}

// Navigate::recalculate (actions/Navigate.cpp:94-144): per port the cumulated jump probabilities
// [stay, d1, d1+d2, ...] with p_i = prob0/exp(decay*dist0) * exp(decay*dist_i); bridges with both ends on land.
// A small host-side table build (ports << cells), like the reference does it at event time.
int recalcNavigation(qhgb_pop *p) {
    qhgb_pop &q = *p;
    if (!q.navNeedUpdate) return 0;
    const int nPorts = (int)q.hPortCell.size();
    if (q.hAltStale && q.haveAlt) {  // the altitude was interpolated on the device: bring the copy up to date
        q.hAlt.resize(q.nCells);
        CK(cudaMemcpyAsync(q.hAlt.data(), q.alt.p, sizeof(double) * q.nCells, cudaMemcpyDeviceToHost, q.stream));
        CK(cudaStreamSynchronize(q.stream));
        q.hAltStale = false;
    }
    const double decay = q.A("Navigate_decay"), A = q.A("Navigate_prob0") / exp(decay * q.A("Navigate_dist0"));
    std::vector<int> row(q.nCells, -1), ptr(nPorts + 1, 0), dest;
    std::vector<double> cum;
    for (int pt = 0; pt < nPorts; pt++) {
        const int b = q.hPortPtr[pt], e = q.hPortPtr[pt + 1];
        ptr[pt] = (int)dest.size();
        row[q.hPortCell[pt]] = pt;
        double sum = 0;
        for (int k = b; k < e; k++) sum += A * exp(decay * q.hDist[k]);
        if (!(sum < 1)) return fail("[Navigate] probabilities for port [%d] add up to %f", q.hPortCell[pt], sum);
        dest.push_back(-1);
        cum.push_back(1 - sum);
        for (int k = b; k < e; k++) {
            dest.push_back(q.hDestCell[k]);
            cum.push_back(cum.back() + A * exp(decay * q.hDist[k]));
        }
    }
    ptr[nPorts] = (int)dest.size();
    std::vector<int2> br;
    for (size_t k = 0; k + 1 < q.hBridges.size(); k += 2) {
        const int a = q.hBridges[k], b = q.hBridges[k + 1];
        if (!q.hAlt.empty() && q.hAlt[a] > 0 && q.hAlt[b] > 0) br.push_back(make_int2(a, b));
    }
    CK(q.navRow.reserve(q.nCells));  // every entry is overwritten below
    CK(q.navPtr.reserve(ptr.size()));
    CK(q.navDest.reserve(std::max<size_t>(dest.size(), 1)));
    CK(q.navCum.reserve(std::max<size_t>(cum.size(), 1)));
    CK(q.navBridges.reserve(std::max<size_t>(br.size(), 1)));
    CK(cudaMemcpyAsync(q.navRow.p, row.data(), row.size() * sizeof(int), cudaMemcpyHostToDevice, q.stream));
    CK(cudaMemcpyAsync(q.navPtr.p, ptr.data(), ptr.size() * sizeof(int), cudaMemcpyHostToDevice, q.stream));
    if (!dest.empty()) CK(cudaMemcpyAsync(q.navDest.p, dest.data(), dest.size() * sizeof(int), cudaMemcpyHostToDevice, q.stream));
    if (!cum.empty()) CK(cudaMemcpyAsync(q.navCum.p, cum.data(), cum.size() * sizeof(double), cudaMemcpyHostToDevice, q.stream));
    if (!br.empty()) CK(cudaMemcpyAsync(q.navBridges.p, br.data(), br.size() * sizeof(int2), cudaMemcpyHostToDevice, q.stream));
    CK(cudaStreamSynchronize(q.stream));
    q.nCurBridges = (int)br.size();
    q.navReady = true;
    q.navNeedUpdate = false;
    return 0;
}

// ConfinedMove::preLoop (actions/ConfinedMove.cpp:44-78), icosahedral branch (the boundary carries no grid type; the flat-grid
// branch compares squared lon/lat differences instead): the cells within ConfinedMove_r km (great circle,
// utils/geomutils.cpp:311-326, RADIUS_EARTH_KM utils/qhg_consts.h:54-55) of (ConfinedMove_x, ConfinedMove_y).  Done once, on the
// host like the reference does it, with the same expressions -- one byte per cell goes to the device.
int recalcConfined(qhgb_pop *p) {
    qhgb_pop &q = *p;
    if (q.hLon.size() != (size_t)q.nCells || q.hLat.size() != (size_t)q.nCells) return fail("[ConfinedMove] no geography (Longitude / Latitude)");
    const double conv = 3.14159 / 180.0, X = q.A("ConfinedMove_x"), Y = q.A("ConfinedMove_y"), R = q.A("ConfinedMove_r");
    std::vector<uint8_t> ok(q.nCells, 0);
    for (int i = 0; i < q.nCells; i++) {
        const double lo1 = q.hLon[i] * conv, la1 = q.hLat[i] * conv, lo2 = X * conv, la2 = Y * conv;
        const double x1 = cos(lo1) * cos(la1), y1 = sin(lo1) * cos(la1), z1 = sin(la1);
        const double x2 = cos(lo2) * cos(la2), y2 = sin(lo2) * cos(la2), z2 = sin(la2);
        double pr = x1 * x2 + y1 * y2 + z1 * z2;
        if (pr > 1) pr = 1; else if (pr < -1) pr = -1;
        if (6371.3 * acos(pr) < R) ok[i] = 1;
    }
    CK(q.allowed.alloc(q.nCells));
    CK(cudaMemcpyAsync(q.allowed.p, ok.data(), ok.size(), cudaMemcpyHostToDevice, q.stream));
    CK(cudaStreamSynchronize(q.stream));
    q.confReady = true;
    return 0;
}

// MoveStats::preLoop (actions/MoveStats.cpp:107-143) + initializeOccupied (:148-186)
int setupMoveStats(qhgb_pop *p) {
    qhgb_pop &q = *p;
    if (q.hLon.size() != (size_t)q.nCells || q.hLat.size() != (size_t)q.nCells) return fail("[MoveStats] no geography (Longitude / Latitude)");
    if (q.nCells > (1 << MS_CELL_BITS)) return fail("[MoveStats] more than %d cells", 1 << MS_CELL_BITS);
    if (q.active(A_CONFINEDMOVE)) return fail("[MoveStats] together with ConfinedMove is not supported (the order of their finalize() calls decides what MoveStats sees)");
    if (q.sharded) return fail("[MoveStats] runs on the generic path, which sharded populations do not have");
    const int mode = (int)q.A("MoveStats_Mode");
    if (mode < 0 || mode > 2) return fail("[MoveStats] MoveStats_Mode %d", mode);
    const size_t n = (size_t)q.nCells;
    CK(q.msHops.alloc(n)); CK(q.msHopsT.alloc(n)); CK(q.msStepHops.alloc(n));
    CK(q.msDist.alloc(n)); CK(q.msTime.alloc(n)); CK(q.msDistT.alloc(n)); CK(q.msTimeT.alloc(n)); CK(q.msLon.alloc(n)); CK(q.msLat.alloc(n));
    CK(q.msKey.alloc(n)); CK(q.msStepDist.alloc(n)); CK(q.msChanged.alloc(n)); CK(q.msDev.alloc(1));
    CK(cudaMemcpyAsync(q.msLon.p, q.hLon.data(), sizeof(double) * n, cudaMemcpyHostToDevice, q.stream));
    CK(cudaMemcpyAsync(q.msLat.p, q.hLat.data(), sizeof(double) * n, cudaMemcpyHostToDevice, q.stream));
    MoveStatsDev &M = q.msHost;
    M.mode = mode;
    M.hops = q.msHops.p; M.dist = q.msDist.p; M.time = q.msTime.p;
    M.hopsT = q.msHopsT.p; M.distT = q.msDistT.p; M.timeT = q.msTimeT.p;
    M.key = q.msKey.p; M.stepHops = q.msStepHops.p; M.stepDist = q.msStepDist.p; M.changed = q.msChanged.p;
    M.lon = q.msLon.p; M.lat = q.msLat.p;
    CK(cudaMemcpyAsync(q.msDev.p, &M, sizeof(M), cudaMemcpyHostToDevice, q.stream));
    LAUNCH(p, "k_move_stats_init", k_move_stats_init, q.gridFor(q.nCells), 256, M, q.nCells, q.count[q.cur].p);
    CK(cudaStreamSynchronize(q.stream));
    q.msReady = true;
    return 0;
}

int computeWeights(qhgb_pop *p) {
    if (!p->haveAlt) return fail("SingleEvaluator[Alt]: no array with name [Altitude]");
    int g = p->gridFor(p->nCells);
    LAUNCH(p, "k_weights_own", k_weights_own, g, 256, p->nCells, p->alt.p, p->haveIce ? p->ice.p : nullptr, p->poly,
           p->havePoly ? 1 : 0, p->W.p);
    LAUNCH(p, "k_weights_cumulate", k_weights_cumulate, g, 256, p->nCells, p->nbr.p, p->W.p, 1);
    CK(cudaGetLastError());
    return 0;
}

CellEnv cellEnv(qhgb_pop *p) {
    CellEnv E{};
    E.nbr = p->nbr.p; E.nNbr = p->nNbr.p; E.ice = p->haveIce ? p->ice.p : nullptr; E.alt = p->alt.p; E.W = p->W.p; E.B = p->B.p; E.D = p->D.p;
    E.TB = p->TB.p; E.TD = p->TD.p;
    if (p->navReady) {
        E.navRow = p->navRow.p; E.navPtr = p->navPtr.p; E.navDest = p->navDest.p; E.navCum = p->navCum.p;
        E.bridges = p->navBridges.p; E.nBridges = p->nCurBridges; E.bridgeProb = p->A("Navigate_bridge_prob");
    }
    E.allowed = p->confReady ? p->allowed.p : nullptr;
    return E;
}

int resetCellCounters(qhgb_pop *p, bool doVerhulst) {
    qhgb_pop &q = *p;
    LAUNCH(p, "k_cell_init", k_cell_init, q.gridFor(q.cHi() - q.cLo()), 256, q.dstats.p, q.cLo(), q.cHi(), q.count[q.cur].p, q.B.p, q.D.p, q.A("Verhulst_b0"),
           q.A("Verhulst_d0"), q.A("Verhulst_theta"), q.A("Verhulst_K"), q.findKind(A_VERHULSTVARK) ? q.cap.p : nullptr, doVerhulst ? 1 : 0, q.stay.p, q.arrive.p, q.cursor.p,
           q.birthCount.p, q.nFert.p, q.TB.p, q.TD.p);
    CK(cudaGetLastError());
    return 0;
}

int ensureCells(qhgb_pop *p) {
    if (p->cellValid) return 0;
    if (p->nAgents > 0) {
        LAUNCH(p, "k_fill_cells", k_fill_cells, p->gridFor((int64_t)(p->cHi() - p->cLo()) * 32), 256, p->cLo(), p->cHi(), p->cellStart[p->cur].p, p->cell[p->cur].p);
        CK(cudaGetLastError());
    }
    p->cellValid = true;
    return 0;
}

// stand-alone pairing (generic path, and whenever the host asks for the mates between initializeStep and finalizeStep)
int ensurePairing(qhgb_pop *p) {
    qhgb_pop &q = *p;
    if (q.pairingValid) return 0;
    if (ensureCells(p) != 0) return -1;
    const int ga = q.gridFor(q.nAgents);
    AgentArrays a = q.arrays(q.cur);
    if (q.needPair) {
        LAUNCH(p, "k_pair_keys", k_pair_keys, ga, 256, q.dstats.p, a, q.key, q.pkey.p, q.mate.p, q.nFert.p);
        LAUNCH(p, "k_pair_rank", k_pair_rank, ga, 256, q.dstats.p, a, q.cellStart[q.cur].p, q.pkey.p, q.prank.p, q.ranked.p);
        LAUNCH(p, "k_pair_match", k_pair_match, ga, 256, q.dstats.p, a, q.cellStart[q.cur].p, q.nFert.p, q.prank.p, q.ranked.p, q.mate.p);
    } else if (q.nAgents > 0) {
        CK(cudaMemsetAsync(q.mate.p, 0xFF, (size_t)q.nAgents * sizeof(int), q.stream));  // -1: nobody is paired
    }
    CK(cudaGetLastError());
    q.pairingValid = true;
    return 0;
}

int launchScan(qhgb_pop *p) {
    qhgb_pop &q = *p;
    const int cA = q.cLo() & ~7, cHi = q.cHi();  // own cells; the start rounded down for aligned 128-bit accesses
    const int nTiles = std::max(1, (cHi - cA + SCAN_TILE - 1) / SCAN_TILE);
    // (one launch instead of two -- tiles taken by ticket, every block waiting for the sums of the tiles before it -- was built and
    // measured: 20.4 us against 18.9 us for the pair, the waiting costs more than the launch)
    LAUNCH(p, "k_scan_tiles", k_scan_tiles, nTiles, 256, cA, cHi, q.stay.p, q.arrive.p, q.birthCount.p, q.tileSums.p);
    LAUNCH(p, "k_scan_apply", k_scan_apply, nTiles, 256, cA, cHi, nTiles, q.stay.p, q.arrive.p, q.birthCount.p, q.tileSums.p,
           q.cellStart[q.cur ^ 1].p, q.birthBase.p, q.count[q.cur ^ 1].p, q.dstats.p, (int)std::min<int64_t>(q.capacity, 2147483647));
    return 0;
}

// the halo of a sharded run: every cell with a neighbour owned by another rank (tools_ico/EQTileLinks.h:20-24 keeps the same sets
// per tile), plus every cell Navigate can send an agent to (destinations of the sea-ways, ends of the bridges: far jumps cross
// any number of shard boundaries).  Arrival counts are exchanged for these cells only; the list is the same on every rank.
int buildHalo(qhgb_pop *p) {
    const int nranks = p->shRanks;
    auto owner = [&](int c) { return (int)(std::upper_bound(p->cellBegin.begin() + 1, p->cellBegin.end(), c) - (p->cellBegin.begin() + 1)); };
    std::vector<uint8_t> mark(p->nCells, 0);
    for (int c = 0; c < p->nCells; c++) {
        const int oc = owner(c);
        for (int j = 0; j < MAXN; j++) {
            const int d = p->hNbr[(size_t)c * MAXN + j];
            if (d >= 0 && owner(d) != oc) { mark[c] = 1; mark[d] = 1; }
        }
    }
    for (int d : p->hDestCell) mark[d] = 1;
    for (int b : p->hBridges) mark[b] = 1;
    std::vector<int> halo;
    for (int c = 0; c < p->nCells; c++) if (mark[c]) halo.push_back(c);
    p->nHalo = (int)halo.size();
    CK(p->dHalo.alloc(halo.size() + 1));
    CK(p->dHaloBuf.alloc(halo.size() + (size_t)nranks * (nranks + 1)));
    if (!halo.empty()) CK(cudaMemcpyAsync(p->dHalo.p, halo.data(), sizeof(int) * halo.size(), cudaMemcpyHostToDevice, p->stream));
    CK(cudaStreamSynchronize(p->stream));
    return 0;
}

// tut_EnvironCapAlt<Mode>Pop -> MultiEvalModes value, -1 for any other class name
int multiProbeMode(const std::string &cls) {
    static const char *const names[] = {nullptr, "tut_EnvironCapAltAddBlockPop", "tut_EnvironCapAltMulPop", "tut_EnvironCapAltMaxPop",
                                        "tut_EnvironCapAltMaxBlockPop", "tut_EnvironCapAltMinPop"};
    for (int m = 1; m <= 5; m++) if (cls == names[m]) return m;
    return -1;
}

// how long a rank waits for its peers at a cross-GPU barrier, in GPU clocks (QHG_XBARRIER_TIMEOUT_S seconds, default 600)
long long xbarrierTimeout() {
    static long long clocks = 0;
    if (!clocks) {
        const char *e = getenv("QHG_XBARRIER_TIMEOUT_S");
        double s = (e && *e) ? atof(e) : 600.0;
        if (!(s > 0)) s = 600.0;
        clocks = (long long)(s * 2.0e9);
    }
    return clocks;
}

// the host has seen a communication error of a sharded step: report it, and clear the device's flags so that the population
// can be read out (the step itself is lost)
int commFailure(qhgb_pop *p) {
    qhgb_pop &q = *p;
    const int err = q.hstats->commError, nRecv = q.hstats->nRecv;
    LAUNCH(p, "k_clear_halt", k_clear_halt, 1, 1, q.dstats.p);
    cudaStreamSynchronize(q.stream);
    if (err == 1) return fail("a rank did not reach the cross-GPU barrier (exchange %u)", q.xStep);
    return fail("receive buffer too small for the migrants of one step (%d > %d)", nRecv, q.recvCap);
}

// pass 2 comes in several compiled shapes: (agents per window, CTAs per SM, cells per grab, window stages)
#define QHG_SCATTER_VARIANTS(X) \
    X(384, 6, 4, 1) X(256, 8, 4, 1) X(192, 6, 4, 2) X(256, 8, 16, 1) X(256, 8, 12, 1) X(384, 6, 16, 1) X(128, 8, 16, 2)
struct ScatterVariant { int sch, minb, sg, nst; };
ScatterVariant scatterVariant(bool sparse) {
    // (the environment is looked at on every call: an A/B run switches between the variants from one step to the next)
    const char *e = getenv(sparse ? "QHG_SCATTER_SPARSE" : "QHG_SCATTER_DENSE");
    ScatterVariant t{};
    if (e && sscanf(e, "%d,%d,%d,%d", &t.sch, &t.minb, &t.sg, &t.nst) == 4) return t;
    return sparse ? ScatterVariant{256, 8, 16, 1} : ScatterVariant{384, 6, 4, 1};  // measured: profiles/ab_scatter_r02*.txt
}

// does a warp of k_seg_decide take 8 cells per grab (fewer than QHG_SEG_DENSE = 64 agents per cell on average) or 4?
bool segSparse(qhgb_pop *p) {
    static int dense = 0;
    if (!dense) {
        const char *e = getenv("QHG_SEG_DENSE");
        dense = (e && atoi(e) > 0) ? atoi(e) : 64;
    }
    const int64_t cells = std::max<int64_t>(1, p->cHi() - p->cLo());
    return (double)p->nAgents / (double)cells < (double)dense;
}

// can the fast path (qhg_cells.cuh) run this program?  The rarer actions exist on the generic path only; Navigate's far jumps
// are handled when it is the last action of the program.
bool programTiled(qhgb_pop *p, const ActParams &P, bool *useNav) {
    qhgb_pop &q = *p;
    bool tiled = !q.forceGeneric, nav = false;
    for (int k = 0; k < P.nOps; k++) {
        const int op = prog_op(P, k);
        if (op == OP_WEIGHTEDMOVERAND || op == OP_SIGDEATH) tiled = false;
        if (op == OP_CONDWEIGHTEDMOVE && !q.segDecide) tiled = false;  // only the batched decide kernel knows it
        if (op == OP_NAVIGATE) {
            if (q.navFast && k == P.nOps - 1 && !P.confine && q.navReady) nav = true;
            else tiled = false;
        }
    }
    if (q.msReady && q.active(A_MOVESTATS)) tiled = false;  // MoveStats sees every registered move: generic path
    if (useNav) *useNav = tiled && nav;
    return tiled;
}

// decide -> scan -> scatter with the given program; used by finalizeStep, by the GEO event and (with an empty
// program, generic path) to bin freshly uploaded agents by cell.
//   tiled   = the fast path (qhg_cells.cuh, one warp per cell): needs the current buffer binned by cell
//   generic = one thread per agent, global atomics; any order of the input, any cell size
int runPipeline(qhgb_pop *p, const ActParams &P, bool advanceStep, bool binned, bool doPair, bool defer = false) {
    qhgb_pop &q = *p;
    const int n = (int)q.nAgents;
    AgentArrays a = q.arrays(q.cur), o = q.arrays(q.cur ^ 1);
    bool useNav = false;  // Navigate on the fast path: it must be the last action, no ConfinedMove
    bool tiled = binned && (n > 0 || q.sharded) && programTiled(p, P, &useNav);
    if (q.sharded && binned && !tiled) return fail("a sharded population only runs on the fast path");
    long long stepEndBirths = -1;
    cudaEvent_t t0 = nullptr, t1 = nullptr;  // device time of the whole pipeline, gaps between the launches included
    if (q.timing) { cudaEventCreate(&t0); cudaEventCreate(&t1); cudaEventRecord(t0, q.stream); }
    for (int attempt = 0; attempt < 2; attempt++) {
        bool stepEndFused = false;  // did the step's last kernel do k_step_end's book-keeping itself?
        if (tiled) {
            const int gridC = q.numSMs * DECIDE_CTAS_PER_SM;  // persistent: 32 warps per SM, one warp per cell at a time
            if (useNav) {
                if (!q.jumps.p) {
                    CK(q.jumps.alloc((size_t)std::max<int64_t>(1 << 16, q.capacity / 16)));
                    CK(q.jumpCount.alloc(1));
                }
                CK(cudaMemsetAsync(q.jumpCount.p, 0, sizeof(int), q.stream));
            }
            const int jumpCap = (int)q.jumps.n;
            const bool sparse = segSparse(p);
            // below 32 agents per cell the compile-time program takes sixteen cells per grab (C2: -9 %; the interpreted kernels with
            // Genetics lose 5 % to the registers of the pending slot reservations and stay at eight; QHG_SEG_SB=8 for A/B)
            static const bool sb8 = [] { const char *e = getenv("QHG_SEG_SB"); return e && atoi(e) == 8; }();
            const bool vsparse = sparse && !sb8 && (double)q.nAgents / (double)std::max<int64_t>(1, q.cHi() - q.cLo()) < 32.0;
            // few grabs per warp (the shards of a many-GPU run): the grabs shrink towards the end of the range
            const int shrinkGrabs = (q.cHi() - q.cLo()) < 128 * q.numSMs * 32 ? 1 : 0;
#define QHG_SEG_LAUNCH_X(NAME, SB_, GEN_, NAV_)                                                                                \
    LAUNCH(p, NAME, (k_seg_decide<false, SB_, GEN_, NAV_>), gridC, DCW * 32, q.dstats.p, a, P, cellEnv(p), q.cLo(), q.cHi(),      \
           q.cellStart[q.cur].p, doPair ? 1 : 0, q.stay.p, q.arrive.p, q.birthCount.p, q.dec.p, q.moveBase.p, shrinkGrabs,        \
           GEN_ ? q.father.p : (int *)nullptr, NAV_ ? q.jumps.p : (JumpEntry *)nullptr, NAV_ ? q.jumpCount.p : (int *)nullptr, jumpCap)
            // the recovery variant of both passes (sharded runs: a rank met a cell beyond the default limits in the first attempt)
            const bool big = q.sharded && (attempt == 1 || q.forceBig);
#define QHG_SEG_LAUNCH_BIG(NAME, GEN_, NAV_)                                                                                          \
    do {                                                                                                                              \
        auto kern = k_seg_decide<false, 4, GEN_, NAV_, true>;                                                                          \
        const int bytes = (int)(DCW * sizeof(SegSmem<4, true>));                                                                       \
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));                                            \
        LAUNCH_SMEM(p, NAME, kern, q.numSMs, DCW * 32, bytes, q.dstats.p, a, P, cellEnv(p), q.cLo(), q.cHi(), q.cellStart[q.cur].p,    \
                    doPair ? 1 : 0, q.stay.p, q.arrive.p, q.birthCount.p, q.dec.p, q.moveBase.p, 1, GEN_ ? q.father.p : (int *)nullptr,  \
                    NAV_ ? q.jumps.p : (JumpEntry *)nullptr, NAV_ ? q.jumpCount.p : (int *)nullptr, jumpCap);                          \
    } while (0)
            if (big) {
                if (q.genetic) LAUNCH(p, "k_genome_ctl_reset", k_genome_ctl_reset, 1, 1, q.gctl.p, 0, 1);
                if (q.genetic && useNav) QHG_SEG_LAUNCH_BIG("k_cell_decide_big", true, true);
                else if (q.genetic) QHG_SEG_LAUNCH_BIG("k_cell_decide_big", true, false);
                else if (useNav) QHG_SEG_LAUNCH_BIG("k_cell_decide_big", false, true);
                else QHG_SEG_LAUNCH_BIG("k_cell_decide_big", false, false);
            } else
            if (q.genetic) {  // births carry the father's position, genome handles follow the agents
                LAUNCH(p, "k_genome_ctl_reset", k_genome_ctl_reset, 1, 1, q.gctl.p, 0, 1);
                if (q.segDecide) {
                    if (useNav) { if (sparse) QHG_SEG_LAUNCH_X("k_cell_decide_genetic_nav", 8, true, true); else QHG_SEG_LAUNCH_X("k_cell_decide_genetic_nav", 4, true, true); }
                    else { if (sparse) QHG_SEG_LAUNCH_X("k_cell_decide_genetic", 8, true, false); else QHG_SEG_LAUNCH_X("k_cell_decide_genetic", 4, true, false); }
                } else if (useNav) {
                    LAUNCH(p, "k_cell_decide_genetic_nav", (k_cell_decide<false, true, true>), gridC, DCW * 32, q.dstats.p, a, P, cellEnv(p), q.cLo(), q.cHi(),
                           q.cellStart[q.cur].p, doPair ? 1 : 0, q.stay.p, q.arrive.p, q.birthCount.p, q.dec.p, q.moveBase.p, q.father.p,
                           q.jumps.p, q.jumpCount.p, jumpCap);
                } else {
                    LAUNCH(p, "k_cell_decide_genetic", (k_cell_decide<false, true>), gridC, DCW * 32, q.dstats.p, a, P, cellEnv(p), q.cLo(), q.cHi(),
                           q.cellStart[q.cur].p, doPair ? 1 : 0, q.stay.p, q.arrive.p, q.birthCount.p, q.dec.p, q.moveBase.p, q.father.p);
                }
            } else if (useNav && q.segDecide) {
                if (sparse) QHG_SEG_LAUNCH_X("k_cell_decide_nav", 8, false, true); else QHG_SEG_LAUNCH_X("k_cell_decide_nav", 4, false, true);
            } else if (useNav) {
                LAUNCH(p, "k_cell_decide_nav", (k_cell_decide<false, false, true>), gridC, DCW * 32, q.dstats.p, a, P, cellEnv(p), q.cLo(), q.cHi(),
                       q.cellStart[q.cur].p, doPair ? 1 : 0, q.stay.p, q.arrive.p, q.birthCount.p, q.dec.p, q.moveBase.p, (int *)nullptr,
                       q.jumps.p, q.jumpCount.p, jumpCap);
            } else if (q.segDecide) {
                // one warp per batch of cells (qhg_decide.cuh); the tutorial action order as straight-line code; 8 cells per grab
                // for sparse populations, 4 for dense ones
                const bool spec = P.prog == PROG_TUT5 && P.nOps == 5 && !P.selfMate && !P.confine && !P.storeAge;
#define QHG_SEG_LAUNCH(NAME, SPEC_, SB_)                                                                                      \
    LAUNCH(p, NAME, (k_seg_decide<SPEC_, SB_>), gridC, DCW * 32, q.dstats.p, a, P, cellEnv(p), q.cLo(), q.cHi(),                  \
           q.cellStart[q.cur].p, doPair ? 1 : 0, q.stay.p, q.arrive.p, q.birthCount.p, q.dec.p, q.moveBase.p, shrinkGrabs)
                if (spec && vsparse) QHG_SEG_LAUNCH("k_cell_decide", true, 16);
                else if (spec && sparse) QHG_SEG_LAUNCH("k_cell_decide", true, 8);
                else if (spec) QHG_SEG_LAUNCH("k_cell_decide", true, 4);
                else if (sparse) QHG_SEG_LAUNCH("k_cell_decide_generic", false, 8);
                else QHG_SEG_LAUNCH("k_cell_decide_generic", false, 4);
#undef QHG_SEG_LAUNCH
#undef QHG_SEG_LAUNCH_X
            } else if (P.prog == PROG_TUT5 && P.nOps == 5 && !P.selfMate && !P.confine) {  // the tutorial action order: compile-time specialised kernel
                LAUNCH(p, "k_cell_decide", k_cell_decide<true>, gridC, DCW * 32, q.dstats.p, a, P, cellEnv(p), q.cLo(), q.cHi(),
                       q.cellStart[q.cur].p, doPair ? 1 : 0, q.stay.p, q.arrive.p, q.birthCount.p, q.dec.p, q.moveBase.p);
            } else {
                LAUNCH(p, "k_cell_decide_generic", k_cell_decide<false>, gridC, DCW * 32, q.dstats.p, a, P, cellEnv(p), q.cLo(), q.cHi(),
                       q.cellStart[q.cur].p, doPair ? 1 : 0, q.stay.p, q.arrive.p, q.birthCount.p, q.dec.p, q.moveBase.p);
            }
            ShardArgs H{};
            long long globalBirths = -1;
            int nRecv = 0;
            std::vector<int> sendCnt, recvCnt;
            if (q.sharded && q.p2p) {
                // exchange over peer memory: remote adds of the arrival counts, births announced, cross-GPU barrier; no host sync
                const int R = q.shRanks, parity = (int)(q.xStep & 1u);
                LAUNCH(p, "k_halo_push", k_halo_push, q.gridFor(q.nHalo), 256, q.nHalo, q.dHalo.p, q.dCellBegin.p, q.shRank, R, q.nCells, parity,
                       q.arrive.p, q.remoteBase.p, q.dPeers.p, q.dstats.p);
                // barrier A and the merge of the remote arrivals in one launch
                LAUNCH(p, "k_xbarrier_merge", k_xbarrier_merge, q.gridFor(std::max(q.nHalo, 1)), 256, q.nHalo, q.dHalo.p, q.cellBegin[q.shRank], q.cellBegin[q.shRank + 1],
                       q.nCells, parity, q.dPeers.p, q.shRank, R, q.xStep + 1, q.arrive.p, q.cursor.p, q.dstats.p, xbarrierTimeout());
                H.on = 1; H.rank = q.shRank; H.nranks = R; H.c0 = q.cellBegin[q.shRank]; H.c1 = q.cellBegin[q.shRank + 1];
                H.cellBegin = q.dCellBegin.p; H.p2p = 1; H.recvCap = q.recvCap; H.remoteBase = q.remoteBase.p; H.peers = q.dPeers.p;
                if (q.genetic) { H.pool = q.gpool.p; H.rowWords = 2 * q.gp.nBlocks; }
                globalBirths = -2;
            } else if (q.sharded) {
                // (1) what this rank sends to every other rank, (2) arrivals per halo cell summed over all ranks,
                // (3) everybody learns every count (and the births per rank: newborn ids are global ranks)
                const int R = q.shRanks;
                // the exchange buffer: nHalo arrival counts, then one row of R+1 ints per rank (everybody fills its own row,
                // the sum over the ranks is the gathered table)
                int *const infoAll = q.dHaloBuf.p + q.nHalo;
                CK(cudaMemsetAsync(infoAll, 0, sizeof(int) * R * (R + 1), q.stream));
                LAUNCH(p, "k_halo_gather", k_halo_gather, q.gridFor(q.nHalo), 256, q.nHalo, q.dHalo.p, q.dCellBegin.p, q.shRank, R, q.arrive.p,
                       q.cursor.p, q.dHaloBuf.p, q.dstats.p, infoAll + q.shRank * (R + 1));
                TIMED(p, "nccl_allreduce_halo", NK(g_nccl.AllReduce(q.dHaloBuf.p, q.dHaloBuf.p, (size_t)q.nHalo + (size_t)R * (R + 1), ncclInt32, ncclSum, q.comm, q.stream)));
                CK(cudaMemcpyAsync(q.hAllInfo, infoAll, sizeof(int) * R * (R + 1), cudaMemcpyDeviceToHost, q.stream));
                LAUNCH(p, "k_halo_apply", k_halo_apply, q.gridFor(q.nHalo), 256, q.nHalo, q.dHalo.p, q.cellBegin[q.shRank], q.cellBegin[q.shRank + 1],
                       q.dHaloBuf.p, q.arrive.p);
                CK(cudaStreamSynchronize(q.stream));
                sendCnt.assign(R, 0); recvCnt.assign(R, 0);
                std::vector<int> sendOff(R + 1, 0);
                long long below = 0, total = 0;
                bool peerOversize = false;  // a rank's pass 1 met a cell beyond its limits: it announced -1 births
                for (int r = 0; r < R; r++) peerOversize |= q.hAllInfo[r * (R + 1) + R] < 0;
                if (peerOversize) {
                    if (attempt == 1 || q.forceBig) return fail("a cell is beyond the limits of the recovery kernels too (8192 agents, 2048 births, 4096 ranked fertile females)");
                    LAUNCH(p, "k_clear_halt", k_clear_halt, 1, 1, q.dstats.p);
                    if (resetCellCounters(p, q.doVerhulst) != 0) return -1;
                    q.bigSteps++;
                    continue;
                }
                for (int r = 0; r < R; r++) {
                    sendCnt[r] = q.hAllInfo[q.shRank * (R + 1) + r];
                    recvCnt[r] = q.hAllInfo[r * (R + 1) + q.shRank];
                    sendOff[r + 1] = sendOff[r] + sendCnt[r];
                    nRecv += recvCnt[r];
                    if (r < q.shRank) below += q.hAllInfo[r * (R + 1) + R];
                    total += q.hAllInfo[r * (R + 1) + R];
                }
                globalBirths = total;
                if ((size_t)sendOff[R] > q.sendBuf.n) CK(q.sendBuf.alloc((size_t)sendOff[R] * 2 + 1024));
                if ((size_t)nRecv > q.recvBuf.n) CK(q.recvBuf.alloc((size_t)nRecv * 2 + 1024));
                if (q.genetic) {  // the genome rows travel in a second pair of buffers, indexed like the records
                    const size_t row = 2 * (size_t)q.gp.nBlocks;
                    if (q.sendBuf.n * row > q.sendGenomes.n) CK(q.sendGenomes.alloc(q.sendBuf.n * row));
                    if (q.recvBuf.n * row > q.recvGenomes.n) CK(q.recvGenomes.alloc(q.recvBuf.n * row));
                    H.pool = q.gpool.p; H.rowWords = (int)row; H.sendGenomes = q.sendGenomes.p;
                }
                CK(cudaMemcpyAsync(q.dSendOff.p, sendOff.data(), sizeof(int) * (R + 1), cudaMemcpyHostToDevice, q.stream));
                CK(cudaMemsetAsync(q.dSendCursor.p, 0, sizeof(int) * R, q.stream));
                H.on = 1; H.rank = q.shRank; H.nranks = R; H.c0 = q.cellBegin[q.shRank]; H.c1 = q.cellBegin[q.shRank + 1];
                H.cellBegin = q.dCellBegin.p; H.sendBuf = q.sendBuf.p; H.sendOff = q.dSendOff.p; H.sendCursor = q.dSendCursor.p;
                H.birthOffset = below;
                q.lastSent = sendOff[R];
                q.lastReceived = nRecv;
            }
            launchScan(p);
            if (q.mirror && !defer) mirrorAfterScan(p, q.cur ^ 1);
            // pass 2: windows of 384 agents, 6 CTAs per SM and 4 cells per grab for dense populations; 256, 8 and 16 for sparse ones
#define QHG_SCATTER_ARGS q.dstats.p, a, o, q.cLo(), q.cHi(), q.cellStart[q.cur].p, q.dec.p, q.nbr.p, q.cellStart[q.cur ^ 1].p, q.stay.p, q.arrive.p, \
                         q.moveBase.p, q.birthBase.p, P.t, P.storeAge, P.selfMate, q.key, H
            // (window size, CTAs per SM, cells per grab, stages): by density, QHG_SCATTER_DENSE / QHG_SCATTER_SPARSE choose another
            // of the compiled variants (A/B runs)
            const ScatterVariant sv = scatterVariant(segSparse(p));
            bool launchedS = false;
#define QHG_SCATTER_CASE(SCH_, MINB_, SG_, NST_)                                                                                            \
    if (!launchedS && sv.sch == SCH_ && sv.minb == MINB_ && sv.sg == SG_ && sv.nst == NST_) {                                              \
        launchedS = true;                                                                                                               \
        if (q.genetic)                                                                                                                  \
            LAUNCH(p, "k_cell_scatter_genetic", (k_cell_scatter<true, SCH_, MINB_, SG_, NST_>), q.numSMs * MINB_, CW * 32, QHG_SCATTER_ARGS,  \
                   q.father.p, q.births.p, q.gctl.p, q.dec.p, shrinkS, 0, 0);                                                           \
        else                                                                                                                            \
            LAUNCH(p, "k_cell_scatter", (k_cell_scatter<false, SCH_, MINB_, SG_, NST_>), q.numSMs * MINB_, CW * 32, QHG_SCATTER_ARGS,         \
                   (const int *)nullptr, (BirthEntry *)nullptr, (GenomeCtl *)nullptr, (uint8_t *)nullptr, shrinkS, seS, advanceStep ? 1 : 0); \
    }
            if (big) {
                launchedS = true;
                if (q.genetic) {
                    auto kern = k_cell_scatter<true, SCH_DENSE, 1, CELL_BATCH, 1, true>;
                    const int bytes = (int)(CW * sizeof(WarpSmemSG<SCH_DENSE, 1, MAXMOTHERS_BIG>));
                    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
                    LAUNCH_SMEM(p, "k_cell_scatter_big", kern, q.numSMs, CW * 32, bytes, QHG_SCATTER_ARGS, q.father.p, q.births.p, q.gctl.p, q.dec.p, 1, 0, 0);
                } else {
                    auto kern = k_cell_scatter<false, SCH_DENSE, 1, CELL_BATCH, 1, true>;
                    const int bytes = (int)(CW * sizeof(WarpSmemS<SCH_DENSE, 1, MAXMOTHERS_BIG>));
                    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
                    LAUNCH_SMEM(p, "k_cell_scatter_big", kern, q.numSMs, CW * 32, bytes, QHG_SCATTER_ARGS, (const int *)nullptr, (BirthEntry *)nullptr,
                                (GenomeCtl *)nullptr, (uint8_t *)nullptr, 1, 0, 0);
                }
            } else {
                // grabs that shrink towards the end of the range: whenever a warp gets fewer than about 24 full grabs
                // without Genetics, Navigate and other ranks pass 2 is the step's last kernel: its last block ends the step
                const int seS = (!q.sharded && !useNav) ? 1 : 0;
                if (seS && !q.genetic) stepEndFused = true;
                const int shrinkS = (shrinkGrabs || (int64_t)(q.cHi() - q.cLo()) < (int64_t)24 * sv.sg * q.numSMs * sv.minb * CW) ? 1 : 0;
                QHG_SCATTER_VARIANTS(QHG_SCATTER_CASE)
            }
#undef QHG_SCATTER_CASE
            if (!launchedS) return fail("no scatter kernel compiled for windows of %d agents, %d CTAs per SM, %d cells per grab, %d stages", sv.sch, sv.minb, sv.sg, sv.nst);
            if (useNav) {
                if (q.genetic) LAUNCH(p, "k_place_jumpers", k_place_jumpers<true>, q.numSMs * 2, 256, q.dstats.p, q.jumpCount.p, q.jumps.p, jumpCap, a, o,
                                      q.cellStart[q.cur ^ 1].p, q.stay.p, q.dec.p, P.storeAge, H);
                else LAUNCH(p, "k_place_jumpers", k_place_jumpers<false>, q.numSMs * 2, 256, q.dstats.p, q.jumpCount.p, q.jumps.p, jumpCap, a, o,
                            q.cellStart[q.cur ^ 1].p, q.stay.p, q.dec.p, P.storeAge, H);
            }
#undef QHG_SCATTER_ARGS
            const int rowW = q.genetic ? 2 * q.gp.nBlocks : 0;
            if (q.sharded && q.p2p) {  // the records are already in the owners' buffers: barrier, then everybody places what it got
                // barrier B inside the placement kernel; without Genetics it is the step's last kernel and ends the step as well
                if (q.genetic) {
                    LAUNCH(p, "k_place_migrants", k_place_migrants_p2p<true>, q.numSMs * 4, 256, q.dstats.p, q.dPeers.p, q.shRank, q.recvCap, o,
                           q.cellStart[q.cur ^ 1].p, q.stay.p, q.cursor.p, P.storeAge, q.gctl.p, q.gfree.p, q.gpool.p, rowW, (int)q.poolRows,
                           q.shRanks, q.xStep + 1, xbarrierTimeout(), 0, 0);
                } else {
                    LAUNCH(p, "k_place_migrants", k_place_migrants_p2p<false>, q.numSMs * 2, 256, q.dstats.p, q.dPeers.p, q.shRank, q.recvCap, o,
                           q.cellStart[q.cur ^ 1].p, q.stay.p, q.cursor.p, P.storeAge, (const GenomeCtl *)nullptr, (const int *)nullptr,
                           (unsigned long long *)nullptr, 0, 0, q.shRanks, q.xStep + 1, xbarrierTimeout(), 1, advanceStep ? 1 : 0);
                    stepEndFused = true;
                }
                q.xStep++;
            } else if (q.sharded) {  // agent migration: packed records between the GPUs (NCCL over NVLink)
                const int R = q.shRanks;
                int so = 0, ro = 0;
                cudaEvent_t g0 = nullptr, g1 = nullptr;
                if (q.timing) { cudaEventCreate(&g0); cudaEventCreate(&g1); cudaEventRecord(g0, q.stream); }
                NK(g_nccl.GroupStart());
                for (int r = 0; r < R; r++) {
                    if (sendCnt[r] > 0) NK(g_nccl.Send(q.sendBuf.p + so, (size_t)sendCnt[r] * sizeof(Migrant), ncclUint8, r, q.comm, q.stream));
                    if (recvCnt[r] > 0) NK(g_nccl.Recv(q.recvBuf.p + ro, (size_t)recvCnt[r] * sizeof(Migrant), ncclUint8, r, q.comm, q.stream));
                    if (q.genetic) {
                        if (sendCnt[r] > 0) NK(g_nccl.Send(q.sendGenomes.p + (size_t)so * rowW, (size_t)sendCnt[r] * rowW, ncclUint64, r, q.comm, q.stream));
                        if (recvCnt[r] > 0) NK(g_nccl.Recv(q.recvGenomes.p + (size_t)ro * rowW, (size_t)recvCnt[r] * rowW, ncclUint64, r, q.comm, q.stream));
                    }
                    so += sendCnt[r];
                    ro += recvCnt[r];
                }
                NK(g_nccl.GroupEnd());
                if (q.timing) { cudaEventRecord(g1, q.stream); q.kt("nccl_sendrecv_migrants").pending.push_back({g0, g1}); }
                if (q.genetic) {
                    LAUNCH(p, "k_place_migrants", k_place_migrants<true>, q.gridFor((int64_t)std::max(nRecv, 1) * 32), 256, q.dstats.p, q.recvBuf.p, nRecv, o,
                           q.cellStart[q.cur ^ 1].p, q.stay.p, q.cursor.p, P.storeAge, q.gctl.p, q.gfree.p, q.gpool.p, q.recvGenomes.p, rowW, (int)q.poolRows);
                } else if (nRecv > 0) {
                    LAUNCH(p, "k_place_migrants", k_place_migrants<false>, q.gridFor(nRecv), 256, q.dstats.p, q.recvBuf.p, nRecv, o,
                           q.cellStart[q.cur ^ 1].p, q.stay.p, q.cursor.p, P.storeAge);
                }
            }
            if (q.genetic) {
                // genomes of the newborns (parents are read from the old buffer); the rows they and the arrivals took are booked;
                // then the rows of the dead and of the agents that left the rank are freed
                LAUNCH(p, "k_make_offspring", k_make_offspring, q.numSMs * 16, 128, q.dstats.p, q.gctl.p, q.births.p, q.gp, q.key,
                       q.gslot[q.cur].p, q.gslot[q.cur ^ 1].p, q.gpool.p, q.gfree.p);
                LAUNCH(p, "k_genome_ctl_reset", k_genome_ctl_reset, 1, 1, q.gctl.p, 1, 0, q.dstats.p, q.sharded ? 1 : 0, (int)q.poolRows);
                LAUNCH(p, "k_free_genomes", k_free_genomes_dec, q.gridFor((n + 3) / 4), 256, q.dstats.p, q.gctl.p, q.dec.p, q.gslot[q.cur].p, q.gfree.p);
            }
            stepEndBirths = globalBirths;
        } else {
            const int ga = q.gridFor(n);
            if (binned && ensureCells(p) != 0) return -1;
            q.needPair = doPair;
            if (ensurePairing(p) != 0) return -1;
            if (q.genetic) LAUNCH(p, "k_genome_ctl_reset", k_genome_ctl_reset, 1, 1, q.gctl.p, 0, 1);
            CellEnv Eg = cellEnv(p);
            const bool doStats = advanceStep && q.msReady && q.active(A_MOVESTATS);  // MoveStats::finalize sees the step's move list
            if (doStats) {
                if (q.nextID >= (1ll << (63 - MS_CELL_BITS))) return fail("[MoveStats] agent ids beyond 2^%d", 63 - MS_CELL_BITS);
                Eg.ms = q.msDev.p;
            }
            LAUNCH(p, "k_actions", k_actions, ga, 256, q.dstats.p, a, q.mate.p, P, Eg, q.arrive.p, q.birthCount.p,
                   q.dest.p, q.rank.p, q.oflags.p);
            if (doStats) {
                LAUNCH(p, "k_move_stats_temp", k_move_stats_temp, q.gridFor(q.nCells), 256, q.msHost, q.nCells, (double)P.t);
                LAUNCH(p, "k_move_stats_merge", k_move_stats_merge, q.gridFor(q.nCells), 256, q.msHost, q.nCells);
            }
            launchScan(p);
            if (q.mirror && !defer) mirrorAfterScan(p, q.cur ^ 1);
            LAUNCH(p, "k_scatter", k_scatter, ga, 256, q.dstats.p, a, o, q.cellStart[q.cur].p, q.dest.p, q.rank.p, q.oflags.p,
                   q.cellStart[q.cur ^ 1].p, q.stay.p, q.arrive.p, q.birthBase.p, P.t, P.storeAge, P.selfMate, q.key, q.mate.p,
                   q.genetic ? q.births.p : nullptr, q.gctl.p);
            if (q.genetic) {  // genomes of the newborns (parents are read from the old buffer), then the rows of the dead are freed
                LAUNCH(p, "k_make_offspring", k_make_offspring, q.numSMs * 16, 128, q.dstats.p, q.gctl.p, q.births.p, q.gp, q.key,
                       q.gslot[q.cur].p, q.gslot[q.cur ^ 1].p, q.gpool.p, q.gfree.p);
                LAUNCH(p, "k_genome_ctl_reset", k_genome_ctl_reset, 1, 1, q.gctl.p, 1, 0);
                LAUNCH(p, "k_free_genomes", k_free_genomes, ga, 256, q.dstats.p, q.gctl.p, q.dest.p, q.gslot[q.cur].p, q.gfree.p);
            }
        }
        if (!stepEndFused) LAUNCH(p, "k_step_end", k_step_end, 1, 1, q.dstats.p, advanceStep ? 1 : 0, stepEndBirths);
        CK(cudaGetLastError());
        if (q.timing && attempt == 0) { cudaEventRecord(t1, q.stream); q.kt("pipeline_total").pending.push_back({t0, t1}); }
        if (defer && tiled) {  // qhgb_run: the host does not wait for the step; the device raises `halt` if it could not complete
            q.tiledSteps++;
            q.cellValid = false;
            q.cur ^= 1;
            q.pairingValid = false;
            return 0;
        }
        // the host's copy of the per-cell counts (m_aiNumAgentsPerCell is current after every step in the reference): queued
        // behind the step, it arrives under the same synchronisation as the step's totals
        if (q.mirror) {
            if (q.mirrorQueued) cudaStreamWaitEvent(q.stream, q.evCopy, 0);
            else enqueueMirror(p, q.cur ^ 1);
            q.mirrorQueued = false;
        }
        if (pullStats(p) != 0) return -1;
        if (q.hstats->commError) return commFailure(p);
        if (tiled && q.sharded && q.p2p) { q.lastSent = q.hstats->nSent; q.lastReceived = q.hstats->nRecv; }
        if (tiled && q.hstats->oversize) {  // a cell too large for the fast path: redo the step on the generic path
            if (q.sharded) {
                // every rank has seen the flag (it travels with the births through the first cross-GPU barrier) and left the step
                // undone: all of them redo it with the recovery kernels
                if (attempt == 1 || q.forceBig) return fail("a cell is beyond the limits of the recovery kernels too (8192 agents, 2048 births, 4096 ranked fertile females)");
                LAUNCH(p, "k_clear_halt", k_clear_halt, 1, 1, q.dstats.p);
                if (resetCellCounters(p, q.doVerhulst) != 0) return -1;
                q.bigSteps++;
                continue;
            }
            tiled = false;
            LAUNCH(p, "k_clear_halt", k_clear_halt, 1, 1, q.dstats.p);
            if (resetCellCounters(p, q.doVerhulst) != 0) return -1;
            continue;
        }
        (tiled ? q.tiledSteps : q.genericSteps)++;
        q.cellValid = !tiled;
        break;
    }
    if (q.hstats->overflow) {
        LAUNCH(p, "k_clear_halt", k_clear_halt, 1, 1, q.dstats.p);  // the step did not happen; the population is as it was
        if (q.mirror) { enqueueMirror(p, q.cur); cudaStreamSynchronize(q.stream); }
        return fail("agent buffers overflowed (capacity %lld, needed %d)", (long long)q.capacity, q.hstats->nNew);
    }
    q.cur ^= 1;
    q.nAgents = q.hstats->nAgents;
    q.nextID = q.hstats->nextID;
    q.agentSteps = q.hstats->agentSteps; q.totSent = q.hstats->totSent; q.totRecv = q.hstats->totRecv;
    q.lastBirths = q.hstats->nBirths;
    q.lastDeaths = q.hstats->nDeaths;
    q.lastMoves = q.hstats->nMoves;
    q.pairingValid = false;
    return 0;
}

}  // namespace

// header and array helpers of qhgb_dump_state / qhgb_restore_state
namespace {
struct DumpHeader {
    char magic[8];          // "QHGB200D"
    uint32_t version;
    char popClass[64];
    int32_t nCells, maxNeigh;
    int64_t nAgents, nextID, stepsDone;
    float curTime;
    uint32_t key[2];
    int32_t genetic, rowWords;
    int32_t evalFirst, evalNeedUpdate, multiFirst, nppNeedUpdate, navNeedUpdate, haveCap, nSubs;
    int32_t subFirst[8], subNeedUpdate[8];
    int64_t lastBirths, lastDeaths, lastMoves;
};
template <class T> bool wr(FILE *f, const std::vector<T> &v) { return v.empty() || fwrite(v.data(), sizeof(T), v.size(), f) == v.size(); }
template <class T> bool rd(FILE *f, std::vector<T> &v) { return v.empty() || fread(v.data(), sizeof(T), v.size(), f) == v.size(); }
}  // namespace

// =================================================================================================
extern "C" {

const char *qhgb_last_error(void) { return g_err.c_str(); }
const char *qhgb_version(void) { return "qhg4_b200 0.1 (sm_100a)"; }

static int create_impl(const char *pop_class, int device, int n_cells, int max_neigh, int64_t capacity_hint, qhgb_pop **out) {
    if (!out) return fail("qhgb_create: out is NULL");
    *out = nullptr;
    if (max_neigh != MAXN) return fail("qhgb_create: connectivity %d not supported (the cell struct holds %d neighbours)", max_neigh, MAXN);
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) return fail("qhgb_create: no CUDA device (%s); there is no CPU fallback", cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail("qhgb_create: device %d of %d", device, ndev);
    CK(cudaSetDevice(device));
    qhgb_pop *p = new qhgb_pop;
    *out = p;  // qhgb_create frees it again if anything below fails
    p->popClass = pop_class;
    p->device = device;
    p->nCells = n_cells;
    p->maxNeigh = max_neigh;
    if (p->popClass == "tut_EnvironAltPop") {  // populations/tut_EnvironAltPop.cpp:24-53
        p->actions = {{"GetOld", A_GETOLD}, {"ATanDeath", A_ATANDEATH}, {"Fertility", A_FERTILITY}, {"Verhulst", A_VERHULST},
                      {"RandomPair", A_RANDOMPAIR}, {"SingleEvaluator[Alt]", A_SINGLEEVAL}, {"WeightedMove", A_WEIGHTEDMOVE}};
    } else if (p->popClass == "tut_EnvironAltNavPop") {
        // tut_EnvironAltPop with Navigate and OldAgeDeath added (actions/Navigate.cpp, actions/OldAgeDeath.cpp): the class the
        // reference driver builds to pin those two actions (NavProbePop, oracle/ref_driver.cpp)
        p->actions = {{"GetOld", A_GETOLD}, {"ATanDeath", A_ATANDEATH}, {"Fertility", A_FERTILITY}, {"Verhulst", A_VERHULST},
                      {"RandomPair", A_RANDOMPAIR}, {"SingleEvaluator[Alt]", A_SINGLEEVAL}, {"WeightedMove", A_WEIGHTEDMOVE},
                      {"Navigate", A_NAVIGATE}, {"OldAgeDeath", A_OLDAGEDEATH}};
    } else if (p->popClass == "tut_EnvironAltConfPop") {
        // tut_EnvironAltPop with ConfinedMove added (actions/ConfinedMove.cpp; carried by 21 of the shipped OoA* classes): the class
        // the reference driver builds to pin the action (ConfProbePop, oracle/ref_driver.cpp)
        p->actions = {{"GetOld", A_GETOLD}, {"ATanDeath", A_ATANDEATH}, {"Fertility", A_FERTILITY}, {"Verhulst", A_VERHULST},
                      {"RandomPair", A_RANDOMPAIR}, {"SingleEvaluator[Alt]", A_SINGLEEVAL}, {"WeightedMove", A_WEIGHTEDMOVE},
                      {"ConfinedMove", A_CONFINEDMOVE}};
    } else if (p->popClass == "tut_ParthenoPop") {  // populations/tut_ParthenoPop.cpp:22-45
        p->actions = {{"GetOld", A_GETOLD}, {"ATanDeath", A_ATANDEATH}, {"RandomMove", A_RANDOMMOVE}, {"Verhulst", A_VERHULST},
                      {"Fertility", A_FERTILITY}};
        p->selfMate = true;
    } else if (p->popClass == "tut_StaticPop") {  // populations/tut_StaticPop.cpp:16-21: no actions
        p->actions = {};
    } else if (p->popClass == "tut_SexualPop") {  // populations/tut_SexualPop.cpp:24-44
        p->actions = {{"GetOld", A_GETOLD}, {"ATanDeath", A_ATANDEATH}, {"RandomMove", A_RANDOMMOVE}, {"Fertility", A_FERTILITY},
                      {"Verhulst", A_VERHULST}, {"RandomPair", A_RANDOMPAIR}};
    } else if (p->popClass == "tut_MovePop") {  // populations/tut_MovePop.cpp:19-31
        p->actions = {{"GetOld", A_GETOLD}, {"ATanDeath", A_ATANDEATH}, {"RandomMove", A_RANDOMMOVE}};
    } else if (p->popClass == "tut_OldAgeDiePop") {  // populations/tut_OldAgeDiePop.cpp:17-26
        p->actions = {{"GetOld", A_GETOLD}, {"ATanDeath", A_ATANDEATH}};
    } else if (p->popClass == "tut_EnvironCapAltPop" || multiProbeMode(p->popClass) >= 0) {  // populations/tut_EnvironCapAltPop.cpp:27-72
        p->actions = {{"GetOld", A_GETOLD}, {"ATanDeath", A_ATANDEATH}, {"Fertility", A_FERTILITY}, {"VerhulstVarK", A_VERHULSTVARK},
                      {"RandomPair", A_RANDOMPAIR}, {"MultiEvaluator[NPP+Alt]", A_MULTIEVAL}, {"WeightedMove", A_WEIGHTEDMOVE},
                      {"NPPCapacity", A_NPPCAP}};
        SubEval ea; ea.input = "Altitude"; ea.weightName = "Multi_weight_alt"; ea.usePoly = true; ea.polyName = "AltPref"; ea.trigger = QHGB_EVENT_ID_GEO;
        SubEval en; en.input = ""; en.weightName = "Multi_weight_npp"; en.usePoly = false; en.trigger = QHGB_EVENT_ID_VEG;
        p->subs = {ea, en};
        if (multiProbeMode(p->popClass) >= 0) {
            // tut_EnvironCapAlt{AddBlock,Mul,Max,MaxBlock,Min}Pop: the same class with its MultiEvaluator combining in another mode
            // over NON-cumulating evaluators and registered as an observer, the way populations/OoANavPop.cpp:50-62 builds its
            // MODE_MUL_SIMPLE evaluator -- the classes the reference driver builds to pin those modes (MultiProbePop<MODE>)
            p->multiMode = multiProbeMode(p->popClass);
            for (auto &e : p->subs) e.cumulate = false;
            p->multiObserves = true;
        }
    } else if (p->popClass == "tut_EnvironAltVarPop") {
        // tut_EnvironAltPop with WeightedMoveRand (actions/WeightedMoveRand.cpp; the predator classes carry it) and SigDeath
        // (actions/SigDeath.cpp) added: the class the reference driver builds to pin them (VarProbePop, oracle/ref_driver.cpp)
        p->actions = {{"GetOld", A_GETOLD}, {"ATanDeath", A_ATANDEATH}, {"Fertility", A_FERTILITY}, {"Verhulst", A_VERHULST},
                      {"RandomPair", A_RANDOMPAIR}, {"SingleEvaluator[Alt]", A_SINGLEEVAL}, {"WeightedMove", A_WEIGHTEDMOVE},
                      {"WeightedMoveRand", A_WEIGHTEDMOVERAND}, {"SigDeath", A_SIGDEATH}};
    } else if (p->popClass.size() == 22 && p->popClass.rfind("tut_EnvironAltCond", 0) == 0 && p->popClass.substr(19) == "Pop" &&
               p->popClass[18] >= '0' && p->popClass[18] <= '7') {
        // tut_EnvironAltPop with CondWeightedMove (actions/CondWeightedMove.cpp; a SimpleCondition of mode m over the altitudes),
        // RandPermPair (actions/RandPermPair.cpp) and MoveStats (actions/MoveStats.cpp) added: the classes the reference driver
        // builds to pin them (ExtProbePop<m>, oracle/ref_driver.cpp)
        p->condMode = p->popClass[18] - '0';
        p->actions = {{"GetOld", A_GETOLD}, {"ATanDeath", A_ATANDEATH}, {"Fertility", A_FERTILITY}, {"Verhulst", A_VERHULST},
                      {"RandomPair", A_RANDOMPAIR}, {"SingleEvaluator[Alt]", A_SINGLEEVAL}, {"WeightedMove", A_WEIGHTEDMOVE},
                      {"CondWeightedMove", A_CONDWEIGHTEDMOVE}, {"RandPermPair", A_RANDPERMPAIR}, {"MoveStats", A_MOVESTATS}};
    } else if (p->popClass == "tut_EnvironAltGenPop" || p->popClass == "tut_EnvironAltGen2bitPop") {
        // tut_EnvironAltPop with Genetics<.., BitGeneUtils> resp. Genetics<.., GeneUtils> added: the classes the reference driver
        // builds to pin the Genetics action (GenProbePop<U>, oracle/ref_driver.cpp)
        p->actions = {{"GetOld", A_GETOLD}, {"ATanDeath", A_ATANDEATH}, {"Fertility", A_FERTILITY}, {"Verhulst", A_VERHULST},
                      {"RandomPair", A_RANDOMPAIR}, {"SingleEvaluator[Alt]", A_SINGLEEVAL}, {"WeightedMove", A_WEIGHTEDMOVE},
                      {"Genetics", A_GENETICS}};
        p->genetic = true;
        p->forceGeneric = true;
        p->gp.bitsPerNuc = (p->popClass == "tut_EnvironAltGen2bitPop") ? 2 : 1;
    } else if (p->popClass == "OoANavGenPop" || p->popClass == "OoANavGen2bitPop") {
        // populations/OoANavGenPop.cpp:33-97; populations/OoANavGen2bitPop.cpp is the same class with Genetics<.., GeneUtils>
        // (2-bit nucleotides, genes/GeneUtils.cpp) and without addObserver(m_pME)
        p->gp.bitsPerNuc = (p->popClass == "OoANavGen2bitPop") ? 2 : 1;
        p->actions = {{"MultiEvaluator[Alt+NPP]", A_MULTIEVAL}, {"WeightedMove", A_WEIGHTEDMOVE}, {"VerhulstVarK", A_VERHULSTVARK},
                      {"RandomPair", A_RANDOMPAIR}, {"GetOld", A_GETOLD}, {"OldAgeDeath", A_OLDAGEDEATH}, {"Fertility", A_FERTILITY},
                      {"NPPCapacity", A_NPPCAP}, {"Genetics", A_GENETICS}, {"Navigate", A_NAVIGATE}};
        SubEval ea; ea.input = "Altitude"; ea.weightName = "Multi_weight_alt"; ea.usePoly = true; ea.polyName = "AltCapPref"; ea.trigger = QHGB_EVENT_ID_GEO;
        SubEval en; en.input = ""; en.weightName = "Multi_weight_npp"; en.usePoly = true; en.polyName = "NPPPref"; en.trigger = QHGB_EVENT_ID_VEG;
        p->subs = {ea, en};
        p->multiObserves = (p->gp.bitsPerNuc == 1);  // addObserver(m_pME), populations/OoANavGenPop.cpp:59; not in OoANavGen2bitPop.cpp
        p->genetic = true;
        p->forceGeneric = true;   // births need the identity of the father: the generic path keeps the full pairing
    } else {
        return fail("qhgb_create: unknown population class [%s]", pop_class);
    }
    // only the classes that override updateEvent drown their agents on EVENT_ID_GEO (populations/tut_EnvironAltPop.cpp:93-127,
    // tut_EnvironCapAltPop.cpp, OoANavGenPop.cpp:179-214; the probe classes derive from tut_EnvironAltPop); the others inherit
    // SPopulation::updateEvent, which does nothing (core/SPopulation.h:116)
    p->drownsOnGeo = p->popClass.rfind("tut_Environ", 0) == 0 || p->popClass.rfind("OoANavGen", 0) == 0;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    p->numSMs = prop.multiProcessorCount;
    CK(cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&p->copyStream, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&p->evScan, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&p->evCopy, cudaEventDisableTiming));
    {
        const char *e = getenv("QHG_B200_PATH");  // "generic" forces the one-thread-per-agent path (testing)
        // populations with Genetics take the fast path too (k_cell_decide<false, true> finds the fathers, k_cell_scatter<true>
        // moves the genome handles and writes the birth records); QHG_GEN_FAST=0 sends them back to the generic path (testing)
        const char *gf = getenv("QHG_GEN_FAST");
        if (p->genetic && !(gf && *gf == '0')) { p->genFast = true; p->forceGeneric = false; }
        // QHG_NAV_FAST=0: programs that end with Navigate go back to the generic path (k_cell_decide<.., true> + k_place_jumpers otherwise)
        const char *dk = getenv("QHG_DECIDE");
        p->segDecide = !(dk && strcmp(dk, "cell") == 0);
        const char *nf = getenv("QHG_NAV_FAST");
        p->navFast = !(nf && *nf == '0');
        p->forceGeneric = p->forceGeneric || (e && strcmp(e, "generic") == 0);
        const char *fb = getenv("QHG_FORCE_BIG");
        p->forceBig = fb && *fb == '1';
    }
    size_t nc = (size_t)n_cells;
    CK(p->nbr.alloc(nc * MAXN));
    CK(p->gid.alloc(nc));
    CK(p->nNbr.alloc(nc));
    CK(p->ice.alloc(nc));
    CK(p->alt.alloc(nc));
    CK(p->count[0].alloc(nc));
    CK(p->count[1].alloc(nc));
    CK(p->moveBase.alloc(nc * MOVE_STRIDE));
    CK(p->count64.alloc(nc));
    CK(p->cellStart[0].alloc(nc + 1));
    CK(p->cellStart[1].alloc(nc + 1));
    CK(p->stay.alloc(nc));
    CK(p->arrive.alloc(nc));
    CK(p->cursor.alloc(nc));
    CK(p->birthCount.alloc(nc));
    CK(p->birthBase.alloc(nc));
    CK(p->nFert.alloc(2 * nc));
    CK(p->W.alloc(nc * WSTRIDE));
    CK(p->B.alloc(nc));
    CK(p->D.alloc(nc));
    CK(p->TB.alloc(nc));
    CK(p->TD.alloc(nc));
    CK(p->tileSums.alloc((nc + SCAN_TILE - 1) / SCAN_TILE + 1));
    if (!p->subs.empty()) {
        CK(p->cap.alloc(nc));
        CK(p->Wtmp.alloc(nc * WSTRIDE));
        CK(cudaMemsetAsync(p->cap.p, 0, nc * sizeof(double), p->stream));
    }
    CK(p->dstats.alloc(1));
    CK(p->gctl.alloc(1));
    CK(cudaMemsetAsync(p->gctl.p, 0, sizeof(GenomeCtl), p->stream));
    CK(cudaMemsetAsync(p->count[0].p, 0, nc * sizeof(int), p->stream));
    CK(cudaMemsetAsync(p->count[1].p, 0, nc * sizeof(int), p->stream));
    CK(cudaMemsetAsync(p->cellStart[0].p, 0, (nc + 1) * sizeof(int), p->stream));
    CK(cudaMemsetAsync(p->cellStart[1].p, 0, (nc + 1) * sizeof(int), p->stream));
    CK(cudaMemsetAsync(p->W.p, 0, nc * WSTRIDE * sizeof(double), p->stream));
    CK(cudaMemsetAsync(p->B.p, 0, nc * sizeof(double), p->stream));
    CK(cudaMemsetAsync(p->D.p, 0, nc * sizeof(double), p->stream));
    CK(cudaMemsetAsync(p->alt.p, 0, nc * sizeof(double), p->stream));
    CK(cudaMemsetAsync(p->ice.p, 0, nc, p->stream));
    CK(cudaMemsetAsync(p->dstats.p, 0, sizeof(DevStats), p->stream));
    CK(cudaMallocHost(&p->hstats, sizeof(DevStats)));
    memset(p->hstats, 0, sizeof(DevStats));
    if (capacity_hint > 0 && allocAgents(p, capacity_hint) != 0) return -1;
    CK(cudaStreamSynchronize(p->stream));
    return 0;
}

int qhgb_destroy(qhgb_pop *p) {
    if (!p) return 0;
    cudaSetDevice(p->device);
    if (p->stream) cudaStreamSynchronize(p->stream);
    for (auto &k : p->ktimes) for (auto &ev : k.pending) { cudaEventDestroy(ev.first); cudaEventDestroy(ev.second); }
    p->nbr.release(); p->gid.release(); p->count[0].release(); p->count[1].release(); p->cellStart[0].release(); p->cellStart[1].release(); p->moveBase.release(); p->count64.release();
    p->stay.release(); p->arrive.release(); p->cursor.release(); p->birthCount.release(); p->birthBase.release(); p->nFert.release(); p->nNbr.release();
    p->ice.release(); p->alt.release(); p->W.release(); p->B.release(); p->D.release(); p->TB.release(); p->TD.release(); p->tileSums.release();
    for (auto &kv : p->envExtra) kv.second.release();
    for (auto &kv : p->envDelta) kv.second.release();
    p->cap.release(); p->Wtmp.release(); p->multiAllowed.release(); p->occCells.release(); p->occOut.release();
    for (int b = 0; b < 2; b++) { p->gslot[b].release(); p->nbabies[b].release(); }
    p->gfree.release(); p->gpool.release(); p->births.release(); p->gctl.release(); p->father.release();
    p->jumps.release(); p->jumpCount.release();
    p->allowed.release();
    p->navRow.release(); p->navPtr.release(); p->navDest.release(); p->navCum.release(); p->navBridges.release();
    for (int b = 0; b < 2; b++) {
        p->id[b].release(); p->birth[b].release(); p->lastBirth[b].release(); p->age[b].release();
        p->cell[b].release(); p->flags[b].release();
    }
    p->mate.release(); p->prank.release(); p->ranked.release(); p->dest.release(); p->rank.release();
    p->oflags.release(); p->dec.release(); p->pkey.release(); p->dstats.release();
    for (void *m : p->ipcOpened) cudaIpcCloseMemHandle(m);
    if (p->arriveRemote) cudaFree(p->arriveRemote);
    if (p->xchg) cudaFree(p->xchg);
    if (p->comm) g_nccl.CommDestroy(p->comm);
    if (p->hAllInfo) cudaFreeHost(p->hAllInfo);
    p->dCellBegin.release(); p->dInfo.release(); p->dAllInfo.release(); p->dSendOff.release(); p->dSendCursor.release();
    p->sendBuf.release(); p->recvBuf.release();
    for (auto &e : p->userEv) if (e) cudaEventDestroy(e);
    if (p->hstats) cudaFreeHost(p->hstats);
    if (p->evScan) cudaEventDestroy(p->evScan);
    if (p->evCopy) cudaEventDestroy(p->evCopy);
    if (p->copyStream) cudaStreamDestroy(p->copyStream);
    if (p->stream) cudaStreamDestroy(p->stream);
    delete p;
    return 0;
}

static int set_cells_impl(qhgb_pop *p, const int32_t *nbr, const int32_t *global_id) {
    if (!p || !nbr) return fail("qhgb_set_cells: NULL argument");
    CK(cudaSetDevice(p->device));
    size_t nc = (size_t)p->nCells;
    std::vector<uint8_t> nn(nc);
    p->hGid.resize(nc);
    for (size_t c = 0; c < nc; c++) {
        int k = 0;
        for (int j = 0; j < MAXN; j++) {
            int v = nbr[c * MAXN + j];
            if (v >= p->nCells) return fail("qhgb_set_cells: cell %zu has neighbour index %d >= %d", c, v, p->nCells);
            // real neighbours come first, -1 pads the tail (core/SCellGrid.cpp:58-62); m_iNumNeighbors = their number
            if (v >= 0) { if (k != j) return fail("qhgb_set_cells: cell %zu has a -1 before a neighbour", c); k++; }
        }
        nn[c] = (uint8_t)k;
        p->hGid[c] = global_id ? global_id[c] : (int32_t)c;
    }
    p->hNbr.assign(nbr, nbr + nc * MAXN);
    CK(cudaMemcpyAsync(p->nbr.p, nbr, nc * MAXN * sizeof(int), cudaMemcpyHostToDevice, p->stream));
    CK(cudaMemcpyAsync(p->nNbr.p, nn.data(), nc, cudaMemcpyHostToDevice, p->stream));
    CK(cudaMemcpyAsync(p->gid.p, p->hGid.data(), nc * sizeof(int), cudaMemcpyHostToDevice, p->stream));
    CK(cudaStreamSynchronize(p->stream));
    p->haveCells = true;
    return 0;
}

static int set_env_array_impl(qhgb_pop *p, const char *name, const double *values, int64_t n) {
    if (!p || !name || !values) return fail("qhgb_set_env_array: NULL argument");
    if (n != p->nCells) return fail("qhgb_set_env_array: [%s] has %lld values, grid has %d cells", name, (long long)n, p->nCells);
    CK(cudaSetDevice(p->device));
    std::string s(name);
    if (s == "Altitude") {
        CK(cudaMemcpyAsync(p->alt.p, values, n * sizeof(double), cudaMemcpyHostToDevice, p->stream));
        p->haveAlt = true;
        p->hAlt.assign(values, values + n);
        p->hAltStale = false;
    } else if (s == "Ice") {
        std::vector<uint8_t> b(n);
        bool any = false;
        for (int64_t i = 0; i < n; i++) { b[i] = values[i] != 0; any |= b[i] != 0; }
        CK(cudaMemcpyAsync(p->ice.p, b.data(), n, cudaMemcpyHostToDevice, p->stream));
        CK(cudaStreamSynchronize(p->stream));
        p->haveIce = true;
        (void)any;
    } else {
        if (s == "Longitude") p->hLon.assign(values, values + n);
        if (s == "Latitude") p->hLat.assign(values, values + n);
        DevBuf<double> &d = p->envExtra[s];
        if (d.n != (size_t)n) CK(d.alloc(n));
        CK(cudaMemcpyAsync(d.p, values, n * sizeof(double), cudaMemcpyHostToDevice, p->stream));
    }
    CK(cudaStreamSynchronize(p->stream));
    return 0;
}

int qhgb_set_env_delta(qhgb_pop *p, const char *name, const double *delta, int64_t n) {
    if (!p || !name) return fail("qhgb_set_env_delta: NULL argument");
    std::string s(name);
    if (!delta) {  // the target is no longer interpolated
        auto it = p->envDelta.find(s);
        if (it != p->envDelta.end()) { it->second.release(); p->envDelta.erase(it); }
        return 0;
    }
    if (n != p->nCells) return fail("qhgb_set_env_delta: [%s] has %lld values, grid has %d cells", name, (long long)n, p->nCells);
    if (s == "Ice") return fail("qhgb_set_env_delta: [Ice] is a flag array and cannot be interpolated");
    if (s == "Altitude" ? !p->haveAlt : !p->envExtra.count(s)) return fail("No array with name [%s] found", name);  // the target must exist
    CK(cudaSetDevice(p->device));
    DevBuf<double> &d = p->envDelta[s];
    if (d.n != (size_t)n) CK(d.alloc(n));
    CK(cudaMemcpyAsync(d.p, delta, n * sizeof(double), cudaMemcpyHostToDevice, p->stream));
    CK(cudaStreamSynchronize(p->stream));
    return 0;
}

int qhgb_get_env_array(qhgb_pop *p, const char *name, double *out) {
    if (!p || !name || !out) return fail("qhgb_get_env_array: NULL argument");
    std::string s(name);
    const double *src = nullptr;
    if (s == "Altitude") src = p->haveAlt ? p->alt.p : nullptr;
    else if (p->envExtra.count(s)) src = p->envExtra[s].p;
    if (!src) return fail("No array with name [%s] found", name);
    CK(cudaSetDevice(p->device));
    CK(cudaMemcpyAsync(out, src, sizeof(double) * (size_t)p->nCells, cudaMemcpyDeviceToHost, p->stream));
    CK(cudaStreamSynchronize(p->stream));
    return 0;
}

int qhgb_interpolate_env(qhgb_pop *p, int steps) {
    if (!p) return fail("qhgb_interpolate_env: NULL population");
    CK(cudaSetDevice(p->device));
    for (auto &kv : p->envDelta) {
        double *target = (kv.first == "Altitude") ? p->alt.p : p->envExtra[kv.first].p;
        LAUNCH(p, "k_env_interpolate", k_env_interpolate, p->gridFor(p->nCells), 256, p->nCells, (double)steps, kv.second.p, target);
        if (kv.first == "Altitude") p->hAltStale = true;
    }
    CK(cudaGetLastError());
    return 0;
}

static const char *const kNumericAttrs[] = {
    "ATanDeath_max_age", "ATanDeath_range", "ATanDeath_slope", "OAD_max_age", "OAD_uncertainty", "WeightedMove_prob", "RandomMove_prob",
    "ConfinedMove_x", "ConfinedMove_y", "ConfinedMove_r", "WeightedMoveRand_prob", "CondWeightedMove_prob", "MoveStats_Mode", "SigDeath_max_age", "SigDeath_range", "SigDeath_slope",
    "Fertility_min_age", "Fertility_max_age", "Fertility_interbirth", "Verhulst_b0", "Verhulst_d0", "Verhulst_theta",
    "Verhulst_K", "NPPCap_water_factor", "NPPCap_coastal_factor", "NPPCap_coastal_min_latitude", "NPPCap_coastal_max_latitude",
    "NPPCap_NPP_min", "NPPCap_NPP_max", "NPPCap_K_max", "NPPCap_K_min", "NPPCap_efficiency", "Multi_weight_alt", "Multi_weight_npp",
    "Navigate_decay", "Navigate_dist0", "Navigate_prob0", "Navigate_min_dens", "Navigate_bridge_prob",
    "Genetics_genome_size", "Genetics_num_crossover", "Genetics_mutation_rate", "Genetics_create_new_genome", "Genetics_bits_per_nuc"};

int qhgb_set_attribute(qhgb_pop *p, const char *name, double value) {
    if (!p || !name) return fail("qhgb_set_attribute: NULL argument");
    for (const char *k : kNumericAttrs) {
        if (strcmp(k, name) == 0) {
            if (strcmp(name, "Genetics_bits_per_nuc") == 0 && (int)value != std::max(1, p->gp.bitsPerNuc))
                return fail("[Genetics] This module expects %d bit nucleotides, but the attribute specifies %d bit nucleotides", std::max(1, p->gp.bitsPerNuc), (int)value);
            if (strcmp(name, "Genetics_genome_size") == 0) {
                if (p->nAgents > 0 && p->genetic) return fail("Genetics_genome_size must be set before agents are added");
                p->gp.genomeSize = (int)value;
                p->gp.nBlocks = ((int)value * std::max(1, p->gp.bitsPerNuc) + 63) / 64;  // numNucs2Blocks, genes/GeneUtils.h:36
                // buffers made from a capacity hint before the row size was known: the genome pool is sized now
                if (p->genetic && p->capacity > 0) {
                    if (cudaSetDevice(p->device) != cudaSuccess) return fail("cudaSetDevice failed");
                    p->poolRows = 0;
                    if (allocAgents(p, p->capacity) != 0) return -1;
                }
            }
            p->attr[name] = value;
            return 0;
        }
    }
    return fail("qhgb_set_attribute: no action of [%s] has an attribute [%s]", p->popClass.c_str(), name);
}

int qhgb_set_attribute_str(qhgb_pop *p, const char *name, const char *value) {
    if (!p || !name || !value) return fail("qhgb_set_attribute_str: NULL argument");
    if (strcmp(name, "Genetics_initial_muts") == 0) { p->attrStr[name] = value; return 0; }  // genomes come from the host (qhgb_set_genomes)
    if (strcmp(name, "AltCapPref") == 0 || strcmp(name, "AltPref") == 0 || strcmp(name, "NPPPref") == 0) {  // PolyLine::readFromString, utils/PolyLine.cpp:92-127
        std::vector<double> d;
        const char *s = value;
        char *e;
        while (true) {
            while (*s == ' ' || *s == '\t') s++;
            if (!*s) break;
            double v = strtod(s, &e);
            if (e == s) return fail("Bad Function def (number format) : [%s]", value);
            d.push_back(v);
            s = e;
        }
        if (d.size() < 4 || d.size() % 2) return fail("[PolyLine::readFromString] Expected non-zero even number of arguments : [%s]", value);
        size_t np = d.size() / 2;
        if (np > (size_t)MAX_POLY) return fail("poly-line [%s] has more than %d points", name, MAX_POLY);
        PolyLineDev pl{};
        pl.nseg = (int)np - 1;
        for (size_t i = 0; i < np; i++) {
            pl.x[i] = d[2 * i];
            pl.v[i] = d[2 * i + 1];
            if (i > 0) pl.a[i - 1] = (pl.v[i] - pl.v[i - 1]) / (pl.x[i] - pl.x[i - 1]);  // utils/PolyLine.h:24-30
        }
        p->polys[name] = pl;
        if (strcmp(name, "NPPPref") != 0) { p->poly = pl; p->havePoly = true; }
        p->attrStr[name] = value;
        p->evalNeedUpdate = true;
        return 0;
    }
    char *e;
    double v = strtod(value, &e);
    if (e == value) return fail("qhgb_set_attribute_str: [%s] = [%s] is not a number", name, value);
    return qhgb_set_attribute(p, name, v);
}

int qhgb_set_prio(qhgb_pop *p, const char *action_name, int prio) {
    if (!p || !action_name) return fail("qhgb_set_prio: NULL argument");
    HostAction *a = p->find(action_name);
    if (!a) return fail("[Prioritizer::setPrio] tried to add non-existing action '%s'", action_name);
    if (prio < 0) return fail("qhgb_set_prio: negative priority for '%s'", action_name);
    a->prio = prio;
    return 0;
}

int qhgb_enable_action(qhgb_pop *p, const char *action_name, int enabled) {
    if (!p || !action_name) return fail("qhgb_enable_action: NULL argument");
    HostAction *a = p->find(action_name);
    if (!a) return fail("qhgb_enable_action: no action '%s'", action_name);
    a->enabled = enabled != 0;
    return 0;
}

int qhgb_set_seed(qhgb_pop *p, const uint32_t *st) {
    if (!p || !st) return fail("qhgb_set_seed: NULL argument");
    memcpy(p->state16, st, sizeof(p->state16));
    p->key.k0 = p->key.k1 = 0;
    for (int j = 0; j < 16; j += 2) { p->key.k0 ^= st[j]; p->key.k1 ^= st[j + 1]; }
    return 0;
}

static int add_agents_impl(qhgb_pop *p, int64_t n, const int32_t *cell, const int64_t *id, const float *birth_time,
                    const uint8_t *gender, const float *age, const float *last_birth, const uint32_t *life_state) {
    if (!p || !cell || !id || !birth_time || !gender) return fail("qhgb_add_agents: NULL argument");
    if (p->genetic && p->gp.nBlocks <= 0) return fail("[Genetics] Genetics_genome_size must be set before agents are added");
    if (p->genetic && p->preLooped) return fail("[Genetics] adding agents after preLoop is not supported for populations with genomes");
    if (p->inStep) return fail("qhgb_add_agents: called inside a step");
    CK(cudaSetDevice(p->device));
    // pack the live ones (readAgentDataQDF drops nothing, but dead records carry no agent: core/SPopulation.cpp:1689-1741)
    std::vector<int32_t> hc; std::vector<int64_t> hi; std::vector<float> hb, hl, ha; std::vector<uint8_t> hf;
    hc.reserve(n); hi.reserve(n); hb.reserve(n); hl.reserve(n); ha.reserve(n); hf.reserve(n);
    for (int64_t j = 0; j < n; j++) {
        uint32_t life = life_state ? life_state[j] : QHGB_LIFE_STATE_ALIVE;
        if (life == QHGB_LIFE_STATE_DEAD) continue;
        if (cell[j] < 0 || cell[j] >= p->nCells) return fail("[addAgent] agent %lld has cellindex %d", (long long)id[j], cell[j]);
        if (gender[j] > 1) return fail("[addAgent] agent %lld has gender %d", (long long)id[j], (int)gender[j]);
        if (id[j] > p->maxID) p->maxID = id[j];
        if (p->sharded && (cell[j] < p->cellBegin[p->shRank] || cell[j] >= p->cellBegin[p->shRank + 1])) continue;  // another rank's cell
        hc.push_back(cell[j]);
        hi.push_back(id[j]);
        hb.push_back(birth_time[j]);
        hl.push_back(last_birth ? last_birth[j] : 0.0f);
        ha.push_back(age ? age[j] : 0.0f);
        hf.push_back((uint8_t)((gender[j] ? F_MALE : 0) | (((life & ~8u) == QHGB_LIFE_STATE_FERTILE) ? F_FERTILE : 0)));
    }
    int64_t m = (int64_t)hc.size();
    if (m == 0) return 0;
    if (materializeAges(p) != 0 || ensureCells(p) != 0) return -1;
    int64_t total = p->nAgents + m;
    if (ensureCapacity(p, total + total / 2 + 1024) != 0) return -1;
    int b = p->cur;
    int64_t off = p->nAgents;
    CK(cudaMemcpyAsync(p->cell[b].p + off, hc.data(), m * sizeof(int), cudaMemcpyHostToDevice, p->stream));
    CK(cudaMemcpyAsync(p->id[b].p + off, hi.data(), m * sizeof(int64_t), cudaMemcpyHostToDevice, p->stream));
    CK(cudaMemcpyAsync(p->birth[b].p + off, hb.data(), m * sizeof(float), cudaMemcpyHostToDevice, p->stream));
    CK(cudaMemcpyAsync(p->lastBirth[b].p + off, hl.data(), m * sizeof(float), cudaMemcpyHostToDevice, p->stream));
    CK(cudaMemcpyAsync(p->age[b].p + off, ha.data(), m * sizeof(float), cudaMemcpyHostToDevice, p->stream));
    CK(cudaMemcpyAsync(p->flags[b].p + off, hf.data(), m, cudaMemcpyHostToDevice, p->stream));
    if (p->genetic) {  // genome rows in upload order; qhgb_set_genomes fills them
        std::vector<int> hs(m);
        for (int64_t j = 0; j < m; j++) hs[j] = (int)(off + j);
        CK(cudaMemcpyAsync(p->gslot[b].p + off, hs.data(), m * sizeof(int), cudaMemcpyHostToDevice, p->stream));
        CK(cudaMemsetAsync(p->nbabies[b].p + off, 0, m * sizeof(int), p->stream));
        CK(cudaMemsetAsync(p->gpool.p + (size_t)off * 2 * p->gp.nBlocks, 0, (size_t)m * 2 * p->gp.nBlocks * sizeof(unsigned long long), p->stream));
    }
    CK(cudaStreamSynchronize(p->stream));
    p->nAgents = total;
    if (p->preLooped) {  // late additions are binned at once
        p->nextID = std::max(p->nextID, p->maxID + 1);
        ActParams P = buildProgram(p, nullptr, p->curTime);
        P.nOps = 0;
        P.prog = 0;
        int rc = pushStats(p);
        if (rc == 0) rc = resetCellCounters(p, false);
        if (rc == 0) rc = runPipeline(p, P, false, false, false);
        return rc;
    }
    return 0;
}

int qhgb_pre_loop(qhgb_pop *p) {
    if (!p) return fail("qhgb_pre_loop: NULL population");
    if (!p->haveCells) return fail("qhgb_pre_loop: no cells (qhgb_set_cells)");
    CK(cudaSetDevice(p->device));
    if (p->capacity == 0 && ensureCapacity(p, 1024) != 0) return -1;
    if (p->sharded) {  // the id base is the maximum over all ranks
        long long *d = nullptr;
        long long h = p->maxID;
        CK(cudaMalloc(&d, sizeof(long long)));
        CK(cudaMemcpyAsync(d, &h, sizeof(h), cudaMemcpyHostToDevice, p->stream));
        NK(g_nccl.AllReduce(d, d, 1, ncclInt64, ncclMax, p->comm, p->stream));
        CK(cudaMemcpyAsync(&h, d, sizeof(h), cudaMemcpyDeviceToHost, p->stream));
        CK(cudaStreamSynchronize(p->stream));
        cudaFree(d);
        p->maxID = h;
    }
    p->nextID = p->maxID + 1;  // IDGen base, app/Simulator.cpp:94-111
    p->stepsDone = 0;
    if (pushStats(p) != 0) return -1;
    // bin the uploaded agents by cell and count them (updateTotal + updateNumAgentsPerCell, core/SPopulation.cpp:287-288):
    // the step pipeline with an empty action list; ages are carried along
    ActParams P = buildProgram(p, nullptr, 0);
    P.nOps = 0;
        P.prog = 0;
    if (resetCellCounters(p, false) != 0) return -1;
    p->doVerhulst = false;
    if (runPipeline(p, P, false, false, false) != 0) return -1;
    if (p->genetic) {  // Genetics::init (actions/Genetics.cpp:196-267): mutation-count table, genome bookkeeping
        p->gp.numCrossOvers = (int)p->A("Genetics_num_crossover");
        p->gp.mutationRate = p->A("Genetics_mutation_rate");
        if (p->gp.numCrossOvers > MAX_CROSS) return fail("[Genetics] more than %d crossovers", MAX_CROSS);
        p->gp.nBino = 0;
        if (p->gp.mutationRate > 0) {
            std::vector<double> t = binomialTable(p->gp.mutationRate, 2 * p->gp.genomeSize, 1e-6);
            if (t.size() > (size_t)MAX_BINO) return fail("[Genetics] mutation-count table has %zu entries (> %d)", t.size(), MAX_BINO);
            p->gp.nBino = (int)t.size();
            for (size_t i = 0; i < t.size(); i++) p->gp.bino[i] = t[i];
        }
        GenomeCtl ctl{0, (int)p->nAgents, 0, 0};
        CK(cudaMemcpyAsync(p->gctl.p, &ctl, sizeof(ctl), cudaMemcpyHostToDevice, p->stream));
        CK(cudaStreamSynchronize(p->stream));
    }
    if (p->findKind(A_NAVIGATE) && p->findKind(A_NAVIGATE)->prio >= 0) {  // Navigate::preLoop, actions/Navigate.cpp:151-174
        if (p->hPortPtr.empty()) return fail("[Navigate] m_pNavigation is NULL! (qhgb_set_navigation)");
        if (recalcNavigation(p) != 0) return -1;
    }
    if (p->findKind(A_NPPCAP) && recalcCapacities(p) != 0) return -1;  // NPPCapacity::preLoop, actions/NPPCapacity.cpp:92-115
    if (p->findKind(A_CONFINEDMOVE) && p->findKind(A_CONFINEDMOVE)->prio >= 0 && recalcConfined(p) != 0) return -1;  // ConfinedMove::preLoop
    if (p->findKind(A_MOVESTATS) && p->findKind(A_MOVESTATS)->prio >= 0 && setupMoveStats(p) != 0) return -1;           // MoveStats::preLoop
    p->preLooped = true;
    p->evalFirst = true;
    return 0;
}

int qhgb_initialize_step(qhgb_pop *p, float t) {
    if (!p) return fail("qhgb_initialize_step: NULL population");
    if (!p->preLooped) return fail("qhgb_initialize_step: preLoop has not run");
    CK(cudaSetDevice(p->device));
    qhgb_pop &q = *p;
    q.curTime = t;
    q.levels.clear();
    q.inStep = true;
    // births need room: at most one baby per paired female (grown here, before the pairing scratch is filled)
    if (ensureCapacity(p, q.nAgents + q.nAgents / 2 + 1024) != 0) return -1;
    // initialize() of every action that has a priority, in order (core/SPopulation.cpp:394-417)
    HostAction *pair = q.find("RandomPair"), *ev = q.find("SingleEvaluator[Alt]");
    bool doVer = q.active(A_VERHULST) || q.active(A_VERHULSTVARK);
    if (q.active(A_VERHULST) && !(q.A("Verhulst_K", 0) != 0)) return fail("Verhulst: Verhulst_K is not set");
    if (resetCellCounters(p, doVer) != 0) return -1;
    // pairing (RandomPair::initialize) is fused into the decide pass of finalizeStep; ensurePairing() runs it
    // stand-alone if the host looks at the mates before that
    q.needPair = pair && pair->prio >= 0 && pair->enabled;
    // RandPermPair (actions/RandPermPair.cpp:99-196) shuffles the larger sex and mates equal places: a uniformly random injection
    // of the smaller sex into the larger one -- the law the rank-by-key pairing of the device implements for RandomPair
    if (q.active(A_RANDPERMPAIR)) q.needPair = true;
    q.doVerhulst = doVer;
    q.pairingValid = false;
    if (ev && ev->prio >= 0 && ev->enabled && (q.evalNeedUpdate || q.evalFirst)) {
        q.evalFirst = false;
        if (computeWeights(p) != 0) return -1;
    }
    if (q.active(A_MULTIEVAL) && computeMultiWeights(p) != 0) return -1;
    CK(cudaGetLastError());
    return 0;
}

int qhgb_do_actions(qhgb_pop *p, unsigned prio, float t) {
    if (!p) return fail("qhgb_do_actions: NULL population");
    if (!p->inStep) return fail("qhgb_do_actions: initializeStep has not run");
    (void)t;
    p->levels.push_back(prio);
    return 0;
}

static int finalizeStepImpl(qhgb_pop *p, bool defer);
int qhgb_finalize_step(qhgb_pop *p) { return finalizeStepImpl(p, false); }

static int finalizeStepImpl(qhgb_pop *p, bool defer) {
    if (!p) return fail("qhgb_finalize_step: NULL population");
    if (!p->inStep) return fail("qhgb_finalize_step: initializeStep has not run");
    CK(cudaSetDevice(p->device));
    qhgb_pop &q = *p;
    ActParams P = buildProgram(p, &q.levels, q.curTime);
    const bool needAge = programNeedsStoredAge(P);
    if (needAge && materializeAges(p) != 0) return -1;
    P.storeAge = needAge ? 1 : 0;
    if (q.nAgents + q.nAgents / 2 + 1024 > q.capacity) return fail("qhgb_finalize_step: agent buffers too small");
    HostAction *ev = q.find("SingleEvaluator[Alt]");
    if (ev && ev->prio >= 0 && ev->enabled) q.evalNeedUpdate = false;  // SingleEvaluator::finalize, :125-130
    if (q.active(A_MULTIEVAL)) for (auto &e : q.subs) e.needUpdate = false;  // MultiEvaluator::finalize, actions/MultiEvaluator.cpp:189-195
    int rc = runPipeline(p, P, true, true, q.needPair, defer);
    q.inStep = false;
    if (rc != 0) return rc;
    q.stepsDone++;
    if (!needAge) { q.ageValid = false; q.lastAgeTime = q.curTime; }
    return 0;
}

static int stepImpl(qhgb_pop *p, float t, bool defer);
int qhgb_step(qhgb_pop *p, float t) { return stepImpl(p, t, false); }

static int stepImpl(qhgb_pop *p, float t, bool defer) {
    if (!p) return fail("qhgb_step: NULL population");
    int rc = qhgb_initialize_step(p, t);
    if (rc != 0) return rc;
    std::vector<unsigned> lv;
    for (auto &a : p->actions) if (a.prio >= 0) lv.push_back((unsigned)a.prio);
    std::sort(lv.begin(), lv.end());
    lv.erase(std::unique(lv.begin(), lv.end()), lv.end());
    for (unsigned l : lv) rc += qhgb_do_actions(p, l, t);
    rc += finalizeStepImpl(p, defer);
    return rc;
}

// can the steps of a run be queued without waiting for each one?  Only the fast path (its grids do not depend on the
// population size), and not the NCCL exchange (the host sizes its messages in every step)
static bool canDefer(qhgb_pop *p) {
    qhgb_pop &q = *p;
    const char *e = getenv("QHG_RUN_SYNC");
    if (e && *e && *e != '0') return false;
    if (q.forceGeneric || q.nAgents <= 0 || !q.preLooped) return false;
    if (q.sharded && !q.p2p) return false;
    const ActParams P = buildProgram(p, nullptr, 0);  // every level of the step
    return programTiled(p, P, nullptr);  // generic-path actions need the host in every step
}

// n steps at t0, t0+1, ...  The steps are queued on the stream in windows of up to 64 without a host round trip in between;
// the host looks at the device's counters once per window.  A step the fast path cannot complete (a cell too large for it,
// agent buffers too small) raises DevStats::halt on the device, the queued steps after it do nothing, and the host redoes that
// step the way qhgb_step would have done it (generic path, larger buffers) and carries on.  The results are those of n
// qhgb_step calls.
int qhgb_run(qhgb_pop *p, float t0, int n_steps) {
    if (!p) return fail("qhgb_run: NULL population");
    CK(cudaSetDevice(p->device));
    qhgb_pop &q = *p;
    int k = 0, rc = 0;
    while (k < n_steps && rc == 0) {
        if (n_steps - k < 2 || !canDefer(p)) {
            rc = qhgb_step(p, t0 + k);
            k++;
            continue;
        }
        const int W = std::min(n_steps - k, 64);
        const int cur0 = q.cur;
        const int64_t steps0 = q.stepsDone, tiled0 = q.tiledSteps;
        const bool ageValid0 = q.ageValid, cellValid0 = q.cellValid;
        const float lastAge0 = q.lastAgeTime;
        const ActParams P0 = buildProgram(p, nullptr, 0);
        const bool needAge = programNeedsStoredAge(P0);
        int w = 0;
        for (; w < W && rc == 0; w++) rc = stepImpl(p, t0 + k + w, true);
        const int hostErr = rc;
        if (q.mirror) enqueueMirror(p, q.cur);  // the host's copy of the per-cell counts: once per window of queued steps
        if (pullStats(p) != 0) return -1;
        if (q.hstats->commError) return commFailure(p);
        const int ok = (int)(q.hstats->step - (unsigned)steps0);  // steps of this window the device completed
        q.nAgents = q.hstats->nAgents;
        q.nextID = q.hstats->nextID;
        q.agentSteps = q.hstats->agentSteps; q.totSent = q.hstats->totSent; q.totRecv = q.hstats->totRecv;
        q.lastBirths = q.hstats->nBirths; q.lastDeaths = q.hstats->nDeaths; q.lastMoves = q.hstats->nMoves;
        if (q.sharded) { q.lastSent = q.hstats->nSent; q.lastReceived = q.hstats->nRecv; }
        if (!q.hstats->halt) {
            if (hostErr != 0) return hostErr;
            k += w;
            continue;
        }
        // step `ok` of the window failed on the device: the host state goes back to where that step starts
        // (sharded: every rank stopped at the same step -- the flag of a cell beyond the limits travels through the first cross-GPU
        // barrier -- and every rank redoes it below; qhgb_step then takes the recovery kernels)
        if (q.sharded && q.hstats->overflow) return fail("a step of a sharded run could not complete on the fast path (agent buffers too small)");
        q.cur = cur0 ^ (ok & 1);
        q.stepsDone = steps0 + ok;
        q.tiledSteps = tiled0 + ok;
        q.cellValid = ok ? false : cellValid0;
        if (!needAge && ok > 0) { q.ageValid = false; q.lastAgeTime = t0 + k + ok - 1; }
        else { q.ageValid = needAge ? true : ageValid0; q.lastAgeTime = lastAge0; }
        q.inStep = false;
        q.pairingValid = false;
        const bool overflow = q.hstats->overflow != 0;
        const int64_t needed = q.hstats->nNew;
        LAUNCH(p, "k_clear_halt", k_clear_halt, 1, 1, q.dstats.p);
        CK(cudaGetLastError());
        if (overflow && ensureCapacity(p, needed + needed / 2 + 1024) != 0) return -1;
        rc = qhgb_step(p, t0 + k + ok);
        k += ok + 1;
    }
    return rc;
}

int qhgb_get_path_counts(qhgb_pop *p, int64_t *fast, int64_t *generic, int64_t *recovery) {
    if (!p) return fail("qhgb_get_path_counts: NULL population");
    if (fast) *fast = p->tiledSteps;
    if (generic) *generic = p->genericSteps;
    if (recovery) *recovery = p->bigSteps;
    return 0;
}

int qhgb_get_run_totals(qhgb_pop *p, int64_t *agent_steps, int64_t *sent, int64_t *received) {
    if (!p) return fail("qhgb_get_run_totals: NULL population");
    if (agent_steps) *agent_steps = p->agentSteps;
    if (sent) *sent = p->totSent;
    if (received) *received = p->totRecv;
    return 0;
}

int qhgb_synchronize(qhgb_pop *p) {
    if (!p) return fail("qhgb_synchronize: NULL population");
    CK(cudaSetDevice(p->device));
    CK(cudaStreamSynchronize(p->stream));
    return 0;
}

int qhgb_update_event(qhgb_pop *p, int event_id, float t) {
    if (!p) return fail("qhgb_update_event: NULL population");
    if (!p->preLooped) return fail("qhgb_update_event: preLoop has not run");
    CK(cudaSetDevice(p->device));
    if (event_id == QHGB_EVENT_ID_GEO && p->drownsOnGeo) {  // populations/tut_EnvironAltPop.cpp:100-127
        ActParams P = buildProgram(p, nullptr, t);
        P.nOps = 1;
        P.prog = OP_DROWN;
        P.storeAge = p->ageValid ? 1 : 0;
        if (resetCellCounters(p, false) != 0) return -1;
        p->doVerhulst = false;
        int rc = runPipeline(p, P, false, true, false);
        if (rc != 0) return rc;
        // notifyObservers(EVENT_ID_GEO): tut_EnvironAltPop never registers its evaluator as an observer (no addObserver in
        // populations/tut_EnvironAltPop.cpp:24-53, unlike populations/OoANavGenPop.cpp:59), so in the reference
        // SingleEvaluator::notify (actions/SingleEvaluator.cpp:332-346) is never reached and the weights of the first
        // step stay in force.  Replicated as is.
        if (p->evaluatorObserves) p->evalNeedUpdate = true;
    }
    // NPPCapacity registers itself as an observer (actions/NPPCapacity.cpp:69) and reacts to GEO, CLIMATE and VEG (:121-131);
    // the MultiEvaluator of tut_EnvironCapAltPop is never registered, so its weights stay as first computed
    if (event_id == QHGB_EVENT_ID_GEO || event_id == QHGB_EVENT_ID_CLIMATE || event_id == QHGB_EVENT_ID_VEG) p->nppNeedUpdate = true;
    // OoANavGenPop registers its MultiEvaluator (populations/OoANavGenPop.cpp:59), which forwards the event to its evaluators
    // (actions/MultiEvaluator.cpp:203-214): each one reacts to its own trigger id
    if (p->multiObserves) for (auto &e : p->subs) if (e.trigger == event_id) e.needUpdate = true;
    if (event_id == QHGB_EVENT_ID_GEO || event_id == QHGB_EVENT_ID_NAV) p->navNeedUpdate = true;  // Navigate::notify, actions/Navigate.cpp:79-87
    return 0;
}

int qhgb_flush_events(qhgb_pop *p, float t) {
    (void)t;
    if (!p) return fail("qhgb_flush_events: NULL population");
    // EVENT_ID_FLUSH: NPPCapacity recalculates now (actions/NPPCapacity.cpp:127-130); evaluators recompute at the next
    // initialize (actions/SingleEvaluator.cpp:335-337)
    if (p->findKind(A_NPPCAP)) {
        CK(cudaSetDevice(p->device));
        if (recalcCapacities(p) != 0) return -1;
    }
    if (p->findKind(A_NAVIGATE) && p->findKind(A_NAVIGATE)->prio >= 0 && !p->hPortPtr.empty()) {
        CK(cudaSetDevice(p->device));
        if (recalcNavigation(p) != 0) return -1;
    }
    return 0;
}

int64_t qhgb_get_num_agents_effective(qhgb_pop *p) { return p ? p->nAgents : -1; }

int qhgb_get_num_agents_array(qhgb_pop *p, uint64_t *out) {
    if (!p || !out) return fail("qhgb_get_num_agents_array: NULL argument");
    if (!p->haveCells) return fail("qhgb_get_num_agents_array: call qhgb_set_cells first");
    CK(cudaSetDevice(p->device));
    // widened to the reference's ulong on the device (cells of other ranks: 0), then one copy into the caller's array --
    // a single DMA when that array is page-locked (qhgb_host_alloc)
    LAUNCH(p, "k_counts_u64", k_counts_u64, p->gridFor(p->nCells), 256, 0, p->nCells, p->cLo(), p->cHi(), p->count[p->cur].p, p->count64.p);
    CK(cudaMemcpyAsync(out, p->count64.p, sizeof(uint64_t) * (size_t)p->nCells, cudaMemcpyDeviceToHost, p->stream));
    CK(cudaStreamSynchronize(p->stream));
    return 0;
}

int qhgb_get_num_agents_range(qhgb_pop *p, int32_t cell_begin, int32_t cell_end, uint64_t *out) {
    if (!p || !out) return fail("qhgb_get_num_agents_range: NULL argument");
    if (!p->haveCells) return fail("qhgb_get_num_agents_range: call qhgb_set_cells first");
    if (cell_begin < 0 || cell_end > p->nCells || cell_begin > cell_end) return fail("qhgb_get_num_agents_range: [%d, %d) is not inside [0, %d)", cell_begin, cell_end, p->nCells);
    if (cell_begin == cell_end) return 0;
    CK(cudaSetDevice(p->device));
    LAUNCH(p, "k_counts_u64", k_counts_u64, p->gridFor(cell_end - cell_begin), 256, cell_begin, cell_end, p->cLo(), p->cHi(), p->count[p->cur].p, p->count64.p);
    CK(cudaMemcpyAsync(out, p->count64.p + cell_begin, sizeof(uint64_t) * (size_t)(cell_end - cell_begin), cudaMemcpyDeviceToHost, p->stream));
    CK(cudaStreamSynchronize(p->stream));
    return 0;
}

int qhgb_mirror_num_agents_array(qhgb_pop *p, uint64_t *host, int32_t cell_begin, int32_t cell_end) {
    if (!p) return fail("qhgb_mirror_num_agents_array: NULL population");
    if (!host) { p->mirror = nullptr; return 0; }
    if (!p->haveCells) return fail("qhgb_mirror_num_agents_array: call qhgb_set_cells first");
    if (cell_begin < 0 || cell_end > p->nCells || cell_begin >= cell_end) return fail("qhgb_mirror_num_agents_array: [%d, %d) is not inside [0, %d)", cell_begin, cell_end, p->nCells);
    CK(cudaSetDevice(p->device));
    p->mirror = host; p->mirrorLo = cell_begin; p->mirrorHi = cell_end;
    enqueueMirror(p, p->cur);
    CK(cudaStreamSynchronize(p->stream));
    return 0;
}

int qhgb_get_move_stats(qhgb_pop *p, int32_t *hops, double *dist, double *time) {
    if (!p || !hops || !dist || !time) return fail("qhgb_get_move_stats: NULL argument");
    if (!p->msReady) return fail("qhgb_get_move_stats: the population has no active MoveStats (or preLoop has not run)");
    CK(cudaSetDevice(p->device));
    const size_t n = (size_t)p->nCells;
    CK(cudaMemcpyAsync(hops, p->msHops.p, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, p->stream));
    CK(cudaMemcpyAsync(dist, p->msDist.p, sizeof(double) * n, cudaMemcpyDeviceToHost, p->stream));
    CK(cudaMemcpyAsync(time, p->msTime.p, sizeof(double) * n, cudaMemcpyDeviceToHost, p->stream));
    CK(cudaStreamSynchronize(p->stream));
    return 0;
}

int qhgb_get_occupied(qhgb_pop *p, int32_t n, const int32_t *cells, uint8_t *out) {
    if (!p || (n > 0 && (!cells || !out))) return fail("qhgb_get_occupied: NULL argument");
    if (n <= 0) return 0;
    for (int i = 0; i < n; i++) if (cells[i] < 0 || cells[i] >= p->nCells) return fail("qhgb_get_occupied: cell index %d", cells[i]);
    CK(cudaSetDevice(p->device));
    if (p->occCells.n < (size_t)n) { CK(p->occCells.alloc(n)); CK(p->occOut.alloc(n)); }
    CK(cudaMemcpyAsync(p->occCells.p, cells, sizeof(int) * (size_t)n, cudaMemcpyHostToDevice, p->stream));
    LAUNCH(p, "k_occupied", k_occupied, p->gridFor(n), 256, n, p->occCells.p, p->cLo(), p->cHi(), p->count[p->cur].p, p->occOut.p);
    CK(cudaMemcpyAsync(out, p->occOut.p, (size_t)n, cudaMemcpyDeviceToHost, p->stream));
    CK(cudaStreamSynchronize(p->stream));
    return 0;
}

void *qhgb_host_alloc(size_t bytes) {
    void *ptr = nullptr;
    if (cudaMallocHost(&ptr, bytes ? bytes : 1) != cudaSuccess) { fail("qhgb_host_alloc: %zu bytes of page-locked memory not available", bytes); return nullptr; }
    return ptr;
}

int qhgb_host_free(void *ptr) {
    if (ptr) CK(cudaFreeHost(ptr));
    return 0;
}

int qhgb_get_step_stats(qhgb_pop *p, qhgb_step_stats *out) {
    if (!p || !out) return fail("qhgb_get_step_stats: NULL argument");
    out->num_agents = p->nAgents;
    out->births = p->lastBirths;
    out->deaths = p->lastDeaths;
    out->moves = p->lastMoves;
    out->next_id = p->nextID;
    out->steps_done = p->stepsDone;
    return 0;
}

static int64_t get_agents_impl(qhgb_pop *p, int64_t cap, int32_t *cell, int32_t *cell_id, int64_t *id, float *birth_time,
                        uint8_t *gender, float *age, float *last_birth, uint32_t *life_state, int64_t *mate_id) {
    if (!p) { fail("qhgb_get_agents: NULL population"); return -1; }
    if (cudaSetDevice(p->device) != cudaSuccess) { fail("cudaSetDevice failed"); return -1; }
    int64_t n = p->nAgents, m = std::min(n, cap);
    if (m <= 0) return n;
    if ((cell || cell_id) && ensureCells(p) != 0) return -1;
    int b = p->cur;
    cudaStream_t s = p->stream;
    std::vector<int32_t> hc;
    std::vector<float> hb;
    std::vector<uint8_t> hf;
    bool err = false;
    auto D2H = [&](void *dst, const void *src, size_t bytes) { err |= cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, s) != cudaSuccess; };
    if (cell || cell_id) { hc.resize(m); D2H(hc.data(), p->cell[b].p, m * sizeof(int)); }
    if (id) D2H(id, p->id[b].p, m * sizeof(int64_t));
    if (birth_time || (age && (!p->ageValid))) { hb.resize(m); D2H(hb.data(), p->birth[b].p, m * sizeof(float)); }
    if (gender || life_state) { hf.resize(m); D2H(hf.data(), p->flags[b].p, m); }
    if (age && !(!p->ageValid)) D2H(age, p->age[b].p, m * sizeof(float));
    if (last_birth) D2H(last_birth, p->lastBirth[b].p, m * sizeof(float));
    if (mate_id) {
        if (p->inStep && ensurePairing(p) == 0) {
            int64_t *tmp = nullptr;
            err |= cudaMalloc(&tmp, m * sizeof(int64_t)) != cudaSuccess;
            if (!err) {
                LAUNCH(p, "k_gather_mate_id", k_gather_mate_id, p->gridFor(m), 256, p->dstats.p, p->id[b].p, p->mate.p, tmp);
                D2H(mate_id, tmp, m * sizeof(int64_t));
                cudaStreamSynchronize(s);
                cudaFree(tmp);
            }
        } else {
            for (int64_t i = 0; i < m; i++) mate_id[i] = -3;
        }
    }
    err |= cudaStreamSynchronize(s) != cudaSuccess;
    if (err) { fail("qhgb_get_agents: device to host copy failed: %s", cudaGetErrorString(cudaGetLastError())); return -1; }
    for (int64_t i = 0; i < m; i++) {
        if (cell) cell[i] = hc[i];
        if (cell_id) cell_id[i] = p->hGid.empty() ? hc[i] : p->hGid[hc[i]];
        if (birth_time) birth_time[i] = hb[i];
        if (gender) gender[i] = hf[i] & F_MALE;
        if (life_state) life_state[i] = (hf[i] & F_FERTILE) ? QHGB_LIFE_STATE_FERTILE : QHGB_LIFE_STATE_ALIVE;
        if (age && (!p->ageValid)) age[i] = p->lastAgeTime - hb[i];
    }
    return n;
}

int qhgb_get_env_weights(qhgb_pop *p, double *out) {
    if (!p || !out) return fail("qhgb_get_env_weights: NULL argument");
    CK(cudaSetDevice(p->device));
    CK(cudaMemcpyAsync(out, p->W.p, (size_t)p->nCells * WSTRIDE * sizeof(double), cudaMemcpyDeviceToHost, p->stream));
    CK(cudaStreamSynchronize(p->stream));
    return 0;
}

int qhgb_get_birth_death_probs(qhgb_pop *p, double *b, double *d) {
    if (!p || !b || !d) return fail("qhgb_get_birth_death_probs: NULL argument");
    CK(cudaSetDevice(p->device));
    CK(cudaMemcpyAsync(b, p->B.p, (size_t)p->nCells * sizeof(double), cudaMemcpyDeviceToHost, p->stream));
    CK(cudaMemcpyAsync(d, p->D.p, (size_t)p->nCells * sizeof(double), cudaMemcpyDeviceToHost, p->stream));
    CK(cudaStreamSynchronize(p->stream));
    return 0;
}

static int set_navigation_impl(qhgb_pop *p, int n_ports, const int32_t *port_cell, const int32_t *port_ptr, const int32_t *dest_cell,
                        const double *dist, int n_bridges, const int32_t *bridges) {
    if (!p || (n_ports > 0 && (!port_cell || !port_ptr || !dest_cell || !dist)) || (n_bridges > 0 && !bridges))
        return fail("qhgb_set_navigation: NULL argument");
    if (!p->findKind(A_NAVIGATE)) return fail("qhgb_set_navigation: population [%s] has no Navigate action", p->popClass.c_str());
    if (n_ports < 0 || n_bridges < 0) return fail("qhgb_set_navigation: negative count");
    if (n_ports == 0) {  // a Navigation group without sea-ways: bridges only
        p->hPortCell.clear();
        p->hPortPtr.assign(1, 0);
    } else {
        p->hPortCell.assign(port_cell, port_cell + n_ports);
        p->hPortPtr.assign(port_ptr, port_ptr + n_ports + 1);
    }
    const int nd = n_ports > 0 ? port_ptr[n_ports] : 0;
    // the reference keeps the destinations of a port in a std::map keyed by cell (core/Navigation.h:13-16): ascending order
    p->hDestCell.clear(); p->hDist.clear();
    for (int pt = 0; pt < n_ports; pt++) {
        if (port_cell[pt] < 0 || port_cell[pt] >= p->nCells) return fail("qhgb_set_navigation: port cell %d", port_cell[pt]);
        std::map<int, double> v;  // a destination given twice keeps the last distance, like the reference's map
        for (int k = port_ptr[pt]; k < port_ptr[pt + 1]; k++) {
            if (dest_cell[k] < 0 || dest_cell[k] >= p->nCells) return fail("qhgb_set_navigation: destination cell %d", dest_cell[k]);
            v[dest_cell[k]] = dist[k];
        }
        p->hPortPtr[pt] = (int)p->hDestCell.size();
        for (auto &x : v) { p->hDestCell.push_back(x.first); p->hDist.push_back(x.second); }
    }
    p->hPortPtr[n_ports] = (int)p->hDestCell.size();
    (void)nd;
    if (n_bridges > 0) p->hBridges.assign(bridges, bridges + 2 * (size_t)n_bridges); else p->hBridges.clear();
    for (int b : p->hBridges) if (b < 0 || b >= p->nCells) return fail("qhgb_set_navigation: bridge cell %d", b);
    p->navNeedUpdate = true;
    if (!p->cellBegin.empty() && p->shRanks > 1) {  // sharded: the destinations join the cells whose arrivals are exchanged
        CK(cudaSetDevice(p->device));
        CK(cudaStreamSynchronize(p->stream));
        if (buildHalo(p) != 0) return -1;
    }
    return 0;
}

int qhgb_set_genomes(qhgb_pop *p, int64_t n, const uint64_t *genomes) {
    if (!p || !genomes) return fail("qhgb_set_genomes: NULL argument");
    if (!p->genetic) return fail("qhgb_set_genomes: population [%s] has no Genetics", p->popClass.c_str());
    if (p->preLooped) return fail("qhgb_set_genomes: must be called before preLoop");
    if (n != p->nAgents) return fail("qhgb_set_genomes: %lld genomes for %lld agents", (long long)n, (long long)p->nAgents);
    CK(cudaSetDevice(p->device));
    CK(cudaMemcpyAsync(p->gpool.p, genomes, (size_t)n * 2 * p->gp.nBlocks * sizeof(unsigned long long), cudaMemcpyHostToDevice, p->stream));
    CK(cudaStreamSynchronize(p->stream));
    return 0;
}

int64_t qhgb_get_genomes(qhgb_pop *p, int64_t cap, uint64_t *genomes, int32_t *num_babies) {
    if (!p) { fail("qhgb_get_genomes: NULL population"); return -1; }
    if (!p->genetic) { fail("qhgb_get_genomes: population [%s] has no Genetics", p->popClass.c_str()); return -1; }
    if (cudaSetDevice(p->device) != cudaSuccess) return -1;
    const int64_t n = p->nAgents, m = std::min(n, cap);
    if (m <= 0) return n;
    const int row = 2 * p->gp.nBlocks;
    if (genomes) {
        unsigned long long *tmp = nullptr;
        if (cudaMalloc(&tmp, (size_t)n * row * sizeof(unsigned long long)) != cudaSuccess) { fail("qhgb_get_genomes: out of device memory"); return -1; }
        LAUNCH(p, "k_gather_genomes", k_gather_genomes, p->gridFor(n * row), 256, p->dstats.p, row, p->gslot[p->cur].p, p->gpool.p, tmp);
        cudaMemcpyAsync(genomes, tmp, (size_t)m * row * sizeof(unsigned long long), cudaMemcpyDeviceToHost, p->stream);
        cudaStreamSynchronize(p->stream);
        cudaFree(tmp);
    }
    if (num_babies) {
        cudaMemcpyAsync(num_babies, p->nbabies[p->cur].p, (size_t)m * sizeof(int), cudaMemcpyDeviceToHost, p->stream);
        cudaStreamSynchronize(p->stream);
    }
    if (cudaGetLastError() != cudaSuccess) { fail("qhgb_get_genomes: device error"); return -1; }
    return n;
}

int qhgb_get_capacities(qhgb_pop *p, double *out) {
    if (!p || !out) return fail("qhgb_get_capacities: NULL argument");
    if (!p->cap.p) return fail("qhgb_get_capacities: population [%s] has no NPPCapacity", p->popClass.c_str());
    CK(cudaSetDevice(p->device));
    CK(cudaMemcpyAsync(out, p->cap.p, (size_t)p->nCells * sizeof(double), cudaMemcpyDeviceToHost, p->stream));
    CK(cudaStreamSynchronize(p->stream));
    return 0;
}

int qhgb_atan_death_prob(qhgb_pop *p, int n, const float *age, double *out) {
    if (!p || !age || !out) return fail("qhgb_atan_death_prob: NULL argument");
    if (n <= 0) return 0;
    CK(cudaSetDevice(p->device));
    ActParams P = buildProgram(p, nullptr, 0);
    float *da = nullptr;
    double *dp = nullptr;
    CK(cudaMalloc(&da, n * sizeof(float)));
    CK(cudaMalloc(&dp, n * sizeof(double)));
    CK(cudaMemcpyAsync(da, age, n * sizeof(float), cudaMemcpyHostToDevice, p->stream));
    LAUNCH(p, "k_atan_prob", k_atan_prob, (n + 255) / 256, 256, n, da, dp, P.atanMaxAge, P.atanSlope, P.atanScale);
    CK(cudaMemcpyAsync(out, dp, n * sizeof(double), cudaMemcpyDeviceToHost, p->stream));
    CK(cudaStreamSynchronize(p->stream));
    cudaFree(da);
    cudaFree(dp);
    return 0;
}

int qhgb_event_record(qhgb_pop *p, int slot) {
    if (!p || slot < 0 || slot >= 8) return fail("qhgb_event_record: bad argument");
    CK(cudaSetDevice(p->device));
    if (!p->userEv[slot]) CK(cudaEventCreate(&p->userEv[slot]));
    CK(cudaEventRecord(p->userEv[slot], p->stream));
    return 0;
}

double qhgb_event_elapsed_ms(qhgb_pop *p, int a, int b) {
    if (!p || a < 0 || a >= 8 || b < 0 || b >= 8 || !p->userEv[a] || !p->userEv[b]) { fail("qhgb_event_elapsed_ms: bad argument"); return -1; }
    cudaSetDevice(p->device);
    if (cudaEventSynchronize(p->userEv[b]) != cudaSuccess) { fail("cudaEventSynchronize failed"); return -1; }
    float ms = 0;
    if (cudaEventElapsedTime(&ms, p->userEv[a], p->userEv[b]) != cudaSuccess) { fail("cudaEventElapsedTime failed"); return -1; }
    return ms;
}

int qhgb_comm_get_unique_id(void *out, int nbytes) {
    if (!out || nbytes < (int)sizeof(ncclUniqueId)) return fail("qhgb_comm_get_unique_id: need a %d byte buffer", (int)sizeof(ncclUniqueId));
    if (loadNccl() != 0) return -1;
    ncclUniqueId id;
    NK(g_nccl.GetUniqueId(&id));
    memcpy(out, &id, sizeof(id));
    return 0;
}

static int comm_init_impl(qhgb_pop *p, int rank, int nranks, const void *unique_id, const int32_t *cell_begin) {
    if (!p || !unique_id || !cell_begin) return fail("qhgb_comm_init: NULL argument");
    if (p->nAgents > 0 || p->preLooped) return fail("qhgb_comm_init: must be called before agents are added");
    if (nranks < 1 || rank < 0 || rank >= nranks) return fail("qhgb_comm_init: rank %d of %d", rank, nranks);
    if (!p->haveCells) return fail("qhgb_comm_init: call qhgb_set_cells first");
    if (cell_begin[0] != 0 || cell_begin[nranks] != p->nCells) return fail("qhgb_comm_init: cell ranges must cover [0, %d)", p->nCells);
    for (int r = 0; r < nranks; r++) if (cell_begin[r + 1] < cell_begin[r]) return fail("qhgb_comm_init: cell ranges must be ascending");
    if (loadNccl() != 0) return -1;
    CK(cudaSetDevice(p->device));
    ncclUniqueId id;
    memcpy(&id, unique_id, sizeof(id));
    NK(g_nccl.CommInitRank(&p->comm, nranks, id, rank));
    p->shRank = rank;
    p->shRanks = nranks;
    p->cellBegin.assign(cell_begin, cell_begin + nranks + 1);
    CK(p->dCellBegin.alloc(nranks + 1));
    CK(p->dInfo.alloc(nranks + 1));
    CK(p->dAllInfo.alloc((size_t)nranks * (nranks + 1)));
    CK(p->dSendOff.alloc(nranks + 1));
    CK(p->dSendCursor.alloc(nranks));
    CK(cudaMallocHost(&p->hAllInfo, sizeof(int) * nranks * (nranks + 1)));
    CK(cudaMemcpyAsync(p->dCellBegin.p, p->cellBegin.data(), sizeof(int) * (nranks + 1), cudaMemcpyHostToDevice, p->stream));
    if (buildHalo(p) != 0) return -1;
    p->sharded = nranks > 1;
    return 0;
}

// ---- dump / restore of the device state -------------------------------------------------------
// The reference dumps a population with SPopulation::dump* / restore* (core/SPopulation.cpp:2024-2570): agent layers,
// the WELL512 states of every thread, the IDGen states.  Here the random streams are counter based (seed, agent id,
// step), so the whole generator state is the step counter; agents are written as the records qhgb_get_agents returns
// (their order inside a cell carries no information).  File: one fixed header, then flat little-endian arrays.

static int dump_state_impl(qhgb_pop *p, const char *path) {
    if (!p || !path) return fail("qhgb_dump_state: NULL argument");
    if (!p->preLooped || p->inStep) return fail("qhgb_dump_state: only between steps (after preLoop / finalizeStep)");
    if (p->subs.size() > 8) return fail("qhgb_dump_state: more than 8 sub-evaluators");
    CK(cudaSetDevice(p->device));
    const int64_t n = p->nAgents;
    std::vector<int32_t> cell(n), cid(n);
    std::vector<int64_t> id(n);
    std::vector<float> birth(n), age(n), last(n);
    std::vector<uint8_t> gender(n);
    std::vector<uint32_t> life(n);
    if (n > 0 && qhgb_get_agents(p, n, cell.data(), cid.data(), id.data(), birth.data(), gender.data(), age.data(), last.data(), life.data(), nullptr) != n) return -1;
    DumpHeader h{};
    memcpy(h.magic, "QHGB200D", 8);
    h.version = 1;
    snprintf(h.popClass, sizeof(h.popClass), "%s", p->popClass.c_str());
    h.nCells = p->nCells; h.maxNeigh = p->maxNeigh; h.nAgents = n; h.nextID = p->nextID; h.stepsDone = p->stepsDone;
    h.curTime = p->curTime; h.key[0] = p->key.k0; h.key[1] = p->key.k1;
    h.genetic = p->genetic ? 1 : 0; h.rowWords = p->genetic ? 2 * p->gp.nBlocks : 0;
    h.evalFirst = p->evalFirst; h.evalNeedUpdate = p->evalNeedUpdate; h.multiFirst = p->multiFirst; h.nppNeedUpdate = p->nppNeedUpdate;
    h.navNeedUpdate = p->navNeedUpdate; h.haveCap = p->cap.p ? 1 : 0; h.nSubs = (int)p->subs.size();
    for (size_t i = 0; i < p->subs.size(); i++) { h.subFirst[i] = p->subs[i].first; h.subNeedUpdate[i] = p->subs[i].needUpdate; }
    h.lastBirths = p->lastBirths; h.lastDeaths = p->lastDeaths; h.lastMoves = p->lastMoves;
    std::vector<uint64_t> genomes((size_t)n * h.rowWords);
    std::vector<int32_t> nbab(p->genetic ? n : 0);
    if (p->genetic && n > 0 && qhgb_get_genomes(p, n, genomes.data(), nbab.data()) != n) return -1;
    std::vector<double> W((size_t)p->nCells * WSTRIDE), cap(h.haveCap ? p->nCells : 0);
    CK(cudaMemcpyAsync(W.data(), p->W.p, W.size() * sizeof(double), cudaMemcpyDeviceToHost, p->stream));
    if (h.haveCap) CK(cudaMemcpyAsync(cap.data(), p->cap.p, cap.size() * sizeof(double), cudaMemcpyDeviceToHost, p->stream));
    CK(cudaStreamSynchronize(p->stream));
    FILE *f = fopen(path, "wb");
    if (!f) return fail("qhgb_dump_state: cannot open [%s]", path);
    bool ok = fwrite(&h, sizeof(h), 1, f) == 1 && wr(f, cell) && wr(f, id) && wr(f, birth) && wr(f, gender) && wr(f, age) && wr(f, last) && wr(f, life) &&
              wr(f, genomes) && wr(f, nbab) && wr(f, W) && wr(f, cap);
    ok = (fclose(f) == 0) && ok;
    if (!ok) return fail("qhgb_dump_state: short write to [%s]", path);
    return 0;
}

static int restore_state_impl(qhgb_pop *p, const char *paths) {
    if (!p || !paths) return fail("qhgb_restore_state: NULL argument");
    if (p->preLooped || p->nAgents > 0) return fail("qhgb_restore_state: the population must be configured (cells, environment, attributes, priorities) but empty");
    // one file, or several separated by newlines: the dumps of ALL ranks of a sharded run.  A sharded population keeps the
    // agents of its own cell range from every file -- which is how a run is re-split over other ranges (or another number of
    // GPUs) after its load has shifted: dump, new ranges, restore.
    std::vector<std::string> files;
    {
        std::string all(paths);
        size_t pos = 0;
        while (pos <= all.size()) {
            const size_t e = all.find('\n', pos);
            const std::string one = all.substr(pos, e == std::string::npos ? std::string::npos : e - pos);
            if (!one.empty()) files.push_back(one);
            if (e == std::string::npos) break;
            pos = e + 1;
        }
    }
    if (files.empty()) return fail("qhgb_restore_state: no file name");
    DumpHeader h{};
    std::vector<int32_t> cell;
    std::vector<int64_t> id;
    std::vector<float> birth, age, last;
    std::vector<uint8_t> gender;
    std::vector<uint32_t> life;
    std::vector<uint64_t> genomes;
    std::vector<int32_t> nbab;
    std::vector<double> W, cap;
    const int c0 = p->cLo(), c1 = p->cHi();
    for (size_t fi = 0; fi < files.size(); fi++) {
        const char *path = files[fi].c_str();
        FILE *f = fopen(path, "rb");
        if (!f) return fail("qhgb_restore_state: cannot open [%s]", path);
        DumpHeader hf{};
        if (fread(&hf, sizeof(hf), 1, f) != 1 || memcmp(hf.magic, "QHGB200D", 8) != 0 || hf.version != 1) { fclose(f); return fail("qhgb_restore_state: [%s] is not a qhg4_b200 dump", path); }
        if (p->popClass != hf.popClass || hf.nCells != p->nCells || hf.maxNeigh != p->maxNeigh || (hf.genetic != 0) != p->genetic ||
            (p->genetic && hf.rowWords != 2 * p->gp.nBlocks) || hf.nSubs != (int)p->subs.size()) {
            fclose(f);
            return fail("qhgb_restore_state: the dump is of [%s], %d cells, %d genome words -- not this population", hf.popClass, hf.nCells, hf.rowWords);
        }
        const int64_t nf = hf.nAgents;
        {   // a corrupt header must not size the buffers below: the counts have to be plausible and the file as long as they say
            if (nf < 0 || nf > (int64_t)2000000000 || hf.rowWords < 0 || hf.rowWords > (1 << 20)) { fclose(f); return fail("qhgb_restore_state: [%s] has a corrupt header (%lld agents)", path, (long long)nf); }
            const long long per = 4 + 8 + 4 + 1 + 4 + 4 + 4 + 8ll * hf.rowWords + (hf.genetic ? 4 : 0);
            const long long need = (long long)sizeof(hf) + nf * per + 8ll * hf.nCells * WSTRIDE + (hf.haveCap ? 8ll * hf.nCells : 0);
            fseek(f, 0, SEEK_END);
            const long long have = ftell(f);
            fseek(f, (long)sizeof(hf), SEEK_SET);
            if (have < need) { fclose(f); return fail("qhgb_restore_state: [%s] is truncated (%lld of %lld bytes)", path, have, need); }
        }
        std::vector<int32_t> fcell(nf);
        std::vector<int64_t> fid(nf);
        std::vector<float> fbirth(nf), fage(nf), flast(nf);
        std::vector<uint8_t> fgender(nf);
        std::vector<uint32_t> flife(nf);
        std::vector<uint64_t> fgen((size_t)nf * hf.rowWords);
        std::vector<int32_t> fnb(hf.genetic ? nf : 0);
        std::vector<double> fW((size_t)hf.nCells * WSTRIDE), fcap(hf.haveCap ? hf.nCells : 0);
        const bool ok = rd(f, fcell) && rd(f, fid) && rd(f, fbirth) && rd(f, fgender) && rd(f, fage) && rd(f, flast) && rd(f, flife) && rd(f, fgen) &&
                        rd(f, fnb) && rd(f, fW) && rd(f, fcap);
        fclose(f);
        if (!ok) return fail("qhgb_restore_state: [%s] is truncated", path);
        if (fi == 0) { h = hf; W.swap(fW); cap.swap(fcap); }
        else if (hf.stepsDone != h.stepsDone) return fail("qhgb_restore_state: [%s] was dumped after %lld steps, the first file after %lld", path, (long long)hf.stepsDone, (long long)h.stepsDone);
        h.nextID = std::max(h.nextID, hf.nextID);
        for (int64_t i = 0; i < nf; i++) {
            if (p->sharded && (fcell[i] < c0 || fcell[i] >= c1)) continue;  // another rank's cell under the ranges in force now
            cell.push_back(fcell[i]); id.push_back(fid[i]); birth.push_back(fbirth[i]); gender.push_back(fgender[i]);
            age.push_back(fage[i]); last.push_back(flast[i]); life.push_back(flife[i]);
            if (hf.genetic) {
                nbab.push_back(fnb[i]);
                genomes.insert(genomes.end(), fgen.begin() + (size_t)i * hf.rowWords, fgen.begin() + (size_t)(i + 1) * hf.rowWords);
            }
        }
    }
    const int64_t n = (int64_t)cell.size();
    CK(cudaSetDevice(p->device));
    p->key.k0 = h.key[0]; p->key.k1 = h.key[1];
    if (n > 0 && qhgb_add_agents(p, n, cell.data(), id.data(), birth.data(), gender.data(), age.data(), last.data(), life.data()) != 0) return -1;
    if (p->genetic && n > 0) {
        if (qhgb_set_genomes(p, n, genomes.data()) != 0) return -1;
        CK(cudaMemcpyAsync(p->nbabies[p->cur].p, nbab.data(), (size_t)n * sizeof(int), cudaMemcpyHostToDevice, p->stream));
        CK(cudaStreamSynchronize(p->stream));
    }
    if (qhgb_pre_loop(p) != 0) return -1;
    // what a fresh start does not know: the generator's step counter, the id base, the weights and capacities as they
    // were (they may lag behind the environment, DESIGN.md §2 "reference quirks"), the observers' flags
    p->stepsDone = h.stepsDone;
    p->nextID = h.nextID;
    p->curTime = h.curTime;
    p->lastBirths = h.lastBirths; p->lastDeaths = h.lastDeaths; p->lastMoves = h.lastMoves;
    if (pushStats(p) != 0) return -1;
    CK(cudaMemcpyAsync(p->W.p, W.data(), W.size() * sizeof(double), cudaMemcpyHostToDevice, p->stream));
    if (h.haveCap && p->cap.p) CK(cudaMemcpyAsync(p->cap.p, cap.data(), cap.size() * sizeof(double), cudaMemcpyHostToDevice, p->stream));
    CK(cudaStreamSynchronize(p->stream));
    p->evalFirst = h.evalFirst != 0; p->evalNeedUpdate = h.evalNeedUpdate != 0; p->multiFirst = h.multiFirst != 0;
    p->nppNeedUpdate = h.nppNeedUpdate != 0; p->navNeedUpdate = h.navNeedUpdate != 0;
    for (size_t i = 0; i < p->subs.size(); i++) { p->subs[i].first = h.subFirst[i] != 0; p->subs[i].needUpdate = h.subNeedUpdate[i] != 0; }
    return 0;
}

int qhgb_comm_p2p_handle(qhgb_pop *p, void *out, int nbytes) {
    if (!p || !out) return fail("qhgb_comm_p2p_handle: NULL argument");
    if (nbytes < 2 * (int)sizeof(cudaIpcMemHandle_t)) return fail("qhgb_comm_p2p_handle: %d bytes needed", 2 * (int)sizeof(cudaIpcMemHandle_t));
    if (p->shRanks < 1 || p->cellBegin.empty()) return fail("qhgb_comm_p2p_handle: call qhgb_comm_init first");
    if (p->shRanks > MAXR) return fail("qhgb_comm_p2p_handle: at most %d ranks", MAXR);
    CK(cudaSetDevice(p->device));
    if (!p->xchg) {
        if (p->genetic && p->gp.nBlocks <= 0) return fail("[Genetics] Genetics_genome_size must be set before the exchange buffers are made");
        p->recvCap = (int)std::max<int64_t>(1 << 16, p->capacity / 8);
        const int rowWords = p->genetic ? 2 * p->gp.nBlocks : 0;
        // header, one record slot per migrant, then (Genetics) one genome row per record slot
        const size_t xb = sizeof(XchgBlock) + (size_t)p->recvCap * (sizeof(Migrant) + (size_t)rowWords * sizeof(unsigned long long));
        CK(cudaMalloc(&p->arriveRemote, sizeof(int) * 2 * (size_t)p->nCells));
        CK(cudaMemset(p->arriveRemote, 0, sizeof(int) * 2 * (size_t)p->nCells));
        CK(cudaMalloc(&p->xchg, xb));
        CK(cudaMemset(p->xchg, 0, sizeof(XchgBlock)));
        XchgBlock hdr{};
        hdr.recvCap = p->recvCap;  // the peers check their slots against the OWNER's capacity
        hdr.rowWords = rowWords;
        CK(cudaMemcpy(p->xchg, &hdr, sizeof(hdr), cudaMemcpyHostToDevice));
        CK(p->remoteBase.alloc((size_t)p->nCells));
    }
    cudaIpcMemHandle_t h[2];
    CK(cudaIpcGetMemHandle(&h[0], p->arriveRemote));
    CK(cudaIpcGetMemHandle(&h[1], p->xchg));
    memcpy(out, h, sizeof(h));
    return 0;
}

int qhgb_comm_p2p_connect(qhgb_pop *p, const void *all_handles) {
    if (!p) return fail("qhgb_comm_p2p_connect: NULL population");
    if (!all_handles) {  // back to the NCCL exchange (e.g. another rank could not map the peers)
        p->p2p = false;
        return 0;
    }
    if (!p->xchg) return fail("qhgb_comm_p2p_connect: call qhgb_comm_p2p_handle first");
    CK(cudaSetDevice(p->device));
    PeerTable T{};
    const cudaIpcMemHandle_t *h = reinterpret_cast<const cudaIpcMemHandle_t *>(all_handles);
    const int myRow = p->genetic ? 2 * p->gp.nBlocks : 0;
    for (int r = 0; r < p->shRanks; r++) {
        if (r == p->shRank) {
            T.arriveRemote[r] = p->arriveRemote;
            T.x[r] = p->xchg;
            T.recvCap[r] = p->recvCap;
            continue;
        }
        void *a = nullptr, *x = nullptr;
        cudaIpcMemHandle_t ha, hx;
        memcpy(&ha, &h[2 * r], sizeof(ha));
        memcpy(&hx, &h[2 * r + 1], sizeof(hx));
        CK(cudaIpcOpenMemHandle(&a, ha, cudaIpcMemLazyEnablePeerAccess));
        p->ipcOpened.push_back(a);
        CK(cudaIpcOpenMemHandle(&x, hx, cudaIpcMemLazyEnablePeerAccess));
        p->ipcOpened.push_back(x);
        T.arriveRemote[r] = (int *)a;
        T.x[r] = (XchgBlock *)x;
        XchgBlock hdr{};  // the owner filled its header before it handed out the handle
        CK(cudaMemcpy(&hdr, x, sizeof(hdr), cudaMemcpyDeviceToHost));
        if (hdr.recvCap <= 0 || hdr.rowWords != myRow) return fail("qhgb_comm_p2p_connect: rank %d has %d record slots and %d-word genome rows (this rank: %d words)", r, hdr.recvCap, hdr.rowWords, myRow);
        T.recvCap[r] = hdr.recvCap;
    }
    CK(p->dPeers.alloc(1));
    CK(cudaMemcpyAsync(p->dPeers.p, &T, sizeof(T), cudaMemcpyHostToDevice, p->stream));
    CK(cudaStreamSynchronize(p->stream));
    p->p2p = p->shRanks > 1;
    return 0;
}

int qhgb_comm_get_traffic(qhgb_pop *p, int64_t *sent, int64_t *received) {
    if (!p) return fail("qhgb_comm_get_traffic: NULL population");
    if (sent) *sent = p->lastSent;
    if (received) *received = p->lastReceived;
    return 0;
}

int64_t qhgb_get_launch_count(qhgb_pop *p) { return p ? p->launches : -1; }
void *qhgb_get_stream(qhgb_pop *p) { return p ? (void *)p->stream : nullptr; }

int qhgb_reset_kernel_times(qhgb_pop *p, int enable) {
    if (!p) return fail("qhgb_reset_kernel_times: NULL population");
    cudaSetDevice(p->device);
    cudaStreamSynchronize(p->stream);
    for (auto &k : p->ktimes) {
        for (auto &ev : k.pending) { cudaEventDestroy(ev.first); cudaEventDestroy(ev.second); }
        k.pending.clear();
        k.ms = 0;
        k.calls = 0;
    }
    p->timing = enable != 0;
    return 0;
}

int qhgb_get_kernel_times(qhgb_pop *p, int cap, const char **names, double *ms, int64_t *calls) {
    if (!p) return fail("qhgb_get_kernel_times: NULL population");
    cudaSetDevice(p->device);
    cudaStreamSynchronize(p->stream);
    for (auto &k : p->ktimes) {
        for (auto &ev : k.pending) {
            float f = 0;
            if (cudaEventElapsedTime(&f, ev.first, ev.second) == cudaSuccess) { k.ms += f; k.calls++; }
            cudaEventDestroy(ev.first);
            cudaEventDestroy(ev.second);
        }
        k.pending.clear();
    }
    int n = (int)p->ktimes.size();
    for (int i = 0; i < n && i < cap; i++) {
        if (names) names[i] = p->ktimes[i].name.c_str();
        if (ms) ms[i] = p->ktimes[i].ms;
        if (calls) calls[i] = p->ktimes[i].calls;
    }
    return n;
}

int qhgb_create(const char *pop_class, int device, int n_cells, int max_neigh, int64_t capacity_hint, qhgb_pop **out) {
    const int rc = guarded([&]() -> int { return create_impl(pop_class, device, n_cells, max_neigh, capacity_hint, out); });
    if (rc != 0 && out && *out) {  // a half-built population is not handed out
        const std::string why = g_err;
        qhgb_destroy(*out);
        *out = nullptr;
        g_err = why;
    }
    return rc;
}

int qhgb_set_cells(qhgb_pop *p, const int32_t *nbr, const int32_t *global_id) {
    return guarded([&]() -> int { return set_cells_impl(p, nbr, global_id); });
}

int qhgb_set_env_array(qhgb_pop *p, const char *name, const double *values, int64_t n) {
    return guarded([&]() -> int { return set_env_array_impl(p, name, values, n); });
}

int qhgb_add_agents(qhgb_pop *p, int64_t n, const int32_t *cell, const int64_t *id, const float *birth_time,
                    const uint8_t *gender, const float *age, const float *last_birth, const uint32_t *life_state) {
    return guarded([&]() -> int { return add_agents_impl(p, n, cell, id, birth_time, gender, age, last_birth, life_state); });
}

int64_t qhgb_get_agents(qhgb_pop *p, int64_t cap, int32_t *cell, int32_t *cell_id, int64_t *id, float *birth_time,
                        uint8_t *gender, float *age, float *last_birth, uint32_t *life_state, int64_t *mate_id) {
    return guarded([&]() -> int64_t { return get_agents_impl(p, cap, cell, cell_id, id, birth_time, gender, age, last_birth, life_state, mate_id); });
}

int qhgb_set_navigation(qhgb_pop *p, int n_ports, const int32_t *port_cell, const int32_t *port_ptr, const int32_t *dest_cell,
                        const double *dist, int n_bridges, const int32_t *bridges) {
    return guarded([&]() -> int { return set_navigation_impl(p, n_ports, port_cell, port_ptr, dest_cell, dist, n_bridges, bridges); });
}

int qhgb_comm_init(qhgb_pop *p, int rank, int nranks, const void *unique_id, const int32_t *cell_begin) {
    return guarded([&]() -> int { return comm_init_impl(p, rank, nranks, unique_id, cell_begin); });
}

int qhgb_dump_state(qhgb_pop *p, const char *path) {
    return guarded([&]() -> int { return dump_state_impl(p, path); });
}

int qhgb_restore_state(qhgb_pop *p, const char *path) {
    return guarded([&]() -> int { return restore_state_impl(p, path); });
}

}  // extern "C"
