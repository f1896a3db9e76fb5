// oracle/ref_driver.cpp -- C driver around the UNMODIFIED reference step loop.
//
// TEST INFRASTRUCTURE ONLY.  This file is compiled by oracle/Makefile together
// with reference sources taken where they lie under /root/reference/QHG4 into
// oracle/_ref/libqhgref.so (git-ignored).  It contains no reference code: it only
// calls the reference's public classes the way app/SimParams.cpp:1553-1590 and
// app/Simulator.cpp:84-111,147-201,287-393 do, minus QDF (HDF5) file I/O:
//   SCellGrid(0,N,{}) + Geography   (core/SCellGrid.cpp:71-84, core/Geography.cpp:21-39)
//   ParamProvider2 -> selectClass -> readSpeciesData   (core/SPopulation.cpp:1108-1142)
//   IDGen::setData(maxID+1, iT, nThreads)               (app/Simulator.cpp:107-111)
//   PopLooper::addPop / preLoop / doStep                (core/PopLooper.cpp:98-113,120-128,166-202)
// Protected members are read for parity checks (compiled with -fno-access-control).
//
// Used by: tests/ (as the checker) and bench.py's cpu_baseline / --impl reference
// leg.  Nothing in the product path links or loads it.
#include <omp.h>
#include <unistd.h>
#include <fcntl.h>
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "hdf5.h"
#include "types.h"
#include "WELL512.h"
#include "PolyLine.h"
#include "SCellGrid.h"
#include "Geography.h"
#include "IDGen.h"
#include "PopLooper.h"
#include "ParamProvider2.h"
#include "ArrayShare.h"
#include "EventConsts.h"
#include "BitGeneUtils.h"
#include "BinomialDist.h"
#include "Climate.h"
#include "Vegetation.h"
#include "tut_EnvironAltPop.h"
#include "tut_EnvironCapAltPop.h"
#include "tut_SexualPop.h"
#include "tut_MovePop.h"
#include "tut_OldAgeDiePop.h"
#include "tut_ParthenoPop.h"
#include "tut_StaticPop.h"
#include "OoANavGenPop.h"  // the genetic population of BASELINE configs #3 / #5, compiled from populations/OoANavGenPop.cpp as it lies
#include "Navigation.h"
#include "Navigate.cpp"  // the reference's action templates, instantiated below for the tutorial agent
#include "OldAgeDeath.cpp"
#include "ConfinedMove.cpp"
#include "WeightedMoveRand.cpp"
#include "CondWeightedMove.cpp"
#include "SimpleCondition.h"
#include "RandPermPair.cpp"
#include "MoveStats.cpp"
#define EPS EPS_SIGDEATH  // actions/SigDeath.h:14 and actions/ATanDeath.h:14 both define a global `EPS`: no reference TU includes both
#include "SigDeath.cpp"
#undef EPS
#include "SingleEvaluator.cpp"  // for the MultiEvaluator probe classes below
#include "MultiEvaluator.cpp"
#include "GeneUtils.h"
#include "LayerArrBuf.cpp"
#include "SequenceIOUtils.cpp"
#include "Genetics.cpp"
#include "GenomeCreator.cpp"
template class SequenceIOUtils<ulong>;  // io/SequenceIOUtils.cpp is an uninstantiated template
#ifdef QHG_WITH_GPU_ADAPTER  // oracle/_ref/libqhgadapter.so: the same driver with the plugin class of INTEGRATION.md in it
#include <cstdlib>
#include <vector>
#include "../integration/tut_EnvironAltGpuPop.h"
#include "../integration/tut_EnvironCapAltGpuPop.h"
#include "../integration/OoANavGenGpuPop.h"
#include "DynPopFactory.h"
#endif

namespace {

struct Quiet {  // the reference prints several lines per step; send them to /dev/null
    int saved = -1;
    explicit Quiet(bool on) {
        if (!on) return;
        fflush(stdout);
        saved = dup(1);
        int nul = open("/dev/null", O_WRONLY);
        dup2(nul, 1);
        close(nul);
    }
    ~Quiet() {
        if (saved < 0) return;
        fflush(stdout);
        dup2(saved, 1);
        close(saved);
    }
};

struct AgentRec {  // the fields every supported population's agent struct has
    int cell; int64_t id; float birth; uint8_t gender; float age; float lastBirth; uint32_t life; int mate; int slot;
};

// what the C API needs from a population, whatever its concrete class
struct PopAccess {
    virtual ~PopAccess() {}
    virtual PopBase *base() = 0;
    virtual int reserve(int n) = 0;
    virtual void put(int slot, const AgentRec &r, gridtype cellID) = 0;
    virtual bool get(int slot, AgentRec &r) = 0;
    virtual int first() = 0;
    virtual int last() = 0;
    virtual idtype &maxID() = 0;
    virtual double *envWeights() = 0;
    virtual double *capacities() = 0;
    virtual int bd(double *&b, double *&d) = 0;
    virtual void atanParams(double &scale, double &slope, double &maxAge) = 0;
    // populations with Genetics: words per genome row (0: none), the row of a slot, the action's own generator (thread 0)
    virtual int genomeWords() { return 0; }
    virtual ulong *genomeRow(int slot) { return nullptr; }
    virtual WELL512 *geneticsWell() { return nullptr; }
    virtual int geneticsInit(int genomeSize, int numCrossOvers, double mutationRate) { return -1; }
    virtual int numBabies(int slot) { return -1; }  // m_iNumBabies of the OoANavGen agents
    virtual ulong *cellCounts() = 0;                // m_aiNumAgentsPerCell
    virtual int moveStats(int *&hops, double *&dist, double *&time) { return -1; }  // the arrays of a MoveStats action
};

template <class PopT, class AgentT>
struct PopAccessT : PopAccess {
    PopT *pop;
    explicit PopAccessT(PopT *p) : pop(p) {}
    PopBase *base() override { return pop; }
    int reserve(int n) override { return pop->reserveAgentSpace(n); }
    void put(int slot, const AgentRec &r, gridtype cellID) override {
        AgentT &a = pop->m_aAgents[slot];
        a.m_iLifeState = r.life; a.m_iCellIndex = r.cell; a.m_ulID = r.id; a.m_ulCellID = cellID;
        a.m_fBirthTime = r.birth; a.m_iGender = r.gender;
        if constexpr (requires { a.m_fAge; }) a.m_fAge = r.age;  // tut_StaticPop keeps the bare Agent
        if constexpr (requires { a.m_fLastBirth; }) { a.m_fLastBirth = r.lastBirth; a.m_iMateIndex = -3; }  // the smaller tutorial agents have no such fields
        if constexpr (requires { a.m_iNumBabies; }) a.m_iNumBabies = 0;  // populations/OoANavGenPop.h:21-27
        // tut_ParthenoPop never writes the mate index of an agent that was read in (populations/tut_ParthenoPop.cpp:76-89);
        // LinearBirth tests it (actions/LinearBirth.cpp:139,142).  In the reference it is whatever LayerBuf's new T[] holds:
        // zero in fresh pages, so the founders do give birth (the tutorial relies on it).  The driver writes that zero.
        if constexpr (std::is_same_v<AgentT, tut_ParthenoAgent>) a.m_iMateIndex = 0;
    }
    bool get(int slot, AgentRec &r) override {
        AgentT &a = pop->m_aAgents[slot];
        if (a.m_iLifeState == LIFE_STATE_DEAD) return false;
        r.cell = a.m_iCellIndex; r.id = a.m_ulID; r.birth = a.m_fBirthTime; r.gender = a.m_iGender; r.age = 0;
        if constexpr (requires { a.m_fAge; }) r.age = a.m_fAge;
        r.lastBirth = 0; r.mate = -3;
        if constexpr (requires { a.m_fLastBirth; }) { r.lastBirth = a.m_fLastBirth; r.mate = a.m_iMateIndex; }
        r.life = a.m_iLifeState; r.slot = slot;
        return true;
    }
    int first() override { return pop->getFirstAgentIndex(); }
    int last() override { return pop->getLastAgentIndex(); }
    idtype &maxID() override { return pop->m_iMaxID; }
    double *envWeights() override {
        if constexpr (requires { pop->m_adEnvWeights; }) return pop->m_adEnvWeights; else return nullptr;
    }
    double *capacities() override {
        if constexpr (requires { pop->m_adCapacities; }) return pop->m_adCapacities; else return nullptr;
    }
    int bd(double *&b, double *&d) override {  // the arrays of LinearBirth / LinearDeath inside Verhulst resp. VerhulstVarK
        if constexpr (requires { pop->m_pVerhulst; }) {
            if (pop->m_pVerhulst->m_pLB == NULL || pop->m_pVerhulst->m_pLD == NULL) return -1;
            b = pop->m_pVerhulst->m_pLB->m_adB; d = pop->m_pVerhulst->m_pLD->m_adD;
            return 0;
        } else if constexpr (requires { pop->m_pVerVarK; }) {
            if (pop->m_pVerVarK->m_pLB == NULL || pop->m_pVerVarK->m_pLD == NULL) return -1;
            b = pop->m_pVerVarK->m_pLB->m_adB; d = pop->m_pVerVarK->m_pLD->m_adD;
            return 0;
        } else {
            return -1;
        }
    }
    int genomeWords() override {
        if constexpr (requires { pop->m_pGenetics; }) return 2 * pop->m_pGenetics->m_iNumBlocks; else return 0;
    }
    ulong *genomeRow(int slot) override {
        if constexpr (requires { pop->m_pGenetics; }) return pop->m_pGenetics->getGenome((uint)slot); else return nullptr;
    }
    // The seed-taking constructor of Genetics never registers its attribute names (actions/Genetics.cpp:91-123, unlike :52-85),
    // so Action::checkAttributes rejects every Genetics attribute of an XML file as unknown: populations with Genetics can
    // only be configured from a QDF (extractAttributesQDF, :440-500, which then calls init()).  Without HDF5 the driver sets
    // the same members and calls the same init().
    int geneticsInit(int genomeSize, int numCrossOvers, double mutationRate) override {
        if constexpr (requires { pop->m_pGenetics; }) {
            pop->m_pGenetics->m_iGenomeSize = genomeSize;
            pop->m_pGenetics->m_iNumCrossOvers = numCrossOvers;
            pop->m_pGenetics->m_dMutationRate = mutationRate;
            pop->m_pGenetics->m_bCreateNewGenome = 0;
            return pop->m_pGenetics->init();
        } else {
            return -1;
        }
    }
    ulong *cellCounts() override { return pop->m_aiNumAgentsPerCell; }
    int moveStats(int *&hops, double *&dist, double *&time) override {
        if constexpr (requires { pop->m_pMS; }) {
            if (pop->m_pMS->m_aiHops == NULL) return -1;
            hops = pop->m_pMS->m_aiHops; dist = pop->m_pMS->m_adDist; time = pop->m_pMS->m_adTime;
            return 0;
        } else {
            return -1;
        }
    }
    int numBabies(int slot) override {
        if constexpr (requires { pop->m_aAgents[slot].m_iNumBabies; }) return pop->m_aAgents[slot].m_iNumBabies; else return -1;
    }
    WELL512 *geneticsWell() override {
        if constexpr (requires { pop->m_pGenetics; }) return pop->m_pGenetics->m_apWELL[0]; else return nullptr;
    }
    void atanParams(double &scale, double &slope, double &maxAge) override {
        if constexpr (requires { pop->m_pAD; }) { scale = pop->m_pAD->m_dScale; slope = pop->m_pAD->m_dSlope; maxAge = pop->m_pAD->m_dMaxAge; }
        else { scale = slope = maxAge = 0; }
    }
};
// Probe class for pinning Navigate and OldAgeDeath (actions/Navigate.cpp, OldAgeDeath.cpp): the reference ships them only inside the large OoA*
// populations (Genetics, QDF sequence I/O).  Here the reference's own Navigate<T> is added to the reference's own
// tut_EnvironAltPop; nothing of the action is restated.
class NavProbePop : public tut_EnvironAltPop {
public:
    NavProbePop(SCellGrid *pCG, PopFinder *pPF, int iLayerSize, IDGen **apIDG, uint32_t *aulState, uint *aiSeeds)
        : tut_EnvironAltPop(pCG, pPF, iLayerSize, apIDG, aulState, aiSeeds) {
        m_pNav = new Navigate<tut_EnvironAltAgent>(this, m_pCG, "", m_apWELL);
        m_prio.addAction(m_pNav);
        m_pOAD = new OldAgeDeath<tut_EnvironAltAgent>(this, m_pCG, "", m_apWELL);  // the other action only the OoA* populations carry
        m_prio.addAction(m_pOAD);
    }
    virtual ~NavProbePop() { delete m_pNav; delete m_pOAD; }
    Navigate<tut_EnvironAltAgent> *m_pNav;
    OldAgeDeath<tut_EnvironAltAgent> *m_pOAD;
};

// Probe class for pinning ConfinedMove (actions/ConfinedMove.cpp:44-101; carried by 21 of the shipped OoA* populations,
// all of which also need Genetics / QDF sequence I/O): the reference's own ConfinedMove<T> added to tut_EnvironAltPop.
class ConfProbePop : public tut_EnvironAltPop {
public:
    ConfProbePop(SCellGrid *pCG, PopFinder *pPF, int iLayerSize, IDGen **apIDG, uint32_t *aulState, uint *aiSeeds)
        : tut_EnvironAltPop(pCG, pPF, iLayerSize, apIDG, aulState, aiSeeds) {
        m_pCM = new ConfinedMove<tut_EnvironAltAgent>(this, m_pCG, "");
        m_prio.addAction(m_pCM);
    }
    virtual ~ConfProbePop() { delete m_pCM; }
    ConfinedMove<tut_EnvironAltAgent> *m_pCM;
};

// Probe class for pinning WeightedMoveRand (actions/WeightedMoveRand.cpp:43-100; carried by the predator populations PDPredPop,
// PDAltPredPop, SimplePredPop, which need a prey population beside them) and SigDeath (actions/SigDeath.cpp:49-90; in no shipped
// class): the reference's own templates added to tut_EnvironAltPop; the parameter file decides by <prio> entries which of
// WeightedMove / WeightedMoveRand and ATanDeath / SigDeath run.
class VarProbePop : public tut_EnvironAltPop {
public:
    VarProbePop(SCellGrid *pCG, PopFinder *pPF, int iLayerSize, IDGen **apIDG, uint32_t *aulState, uint *aiSeeds)
        : tut_EnvironAltPop(pCG, pPF, iLayerSize, apIDG, aulState, aiSeeds) {
        m_pWMR = new WeightedMoveRand<tut_EnvironAltAgent>(this, m_pCG, "", m_apWELL, m_adEnvWeights);
        m_prio.addAction(m_pWMR);
        m_pSD = new SigDeath<tut_EnvironAltAgent>(this, m_pCG, "", m_apWELL);
        m_prio.addAction(m_pSD);
    }
    virtual ~VarProbePop() { delete m_pWMR; delete m_pSD; }
    WeightedMoveRand<tut_EnvironAltAgent> *m_pWMR;
    SigDeath<tut_EnvironAltAgent> *m_pSD;
};

// Probe classes for pinning CondWeightedMove (actions/CondWeightedMove.cpp:41-86, with the reference's only MoveCondition,
// actions/SimpleCondition.cpp, over the altitudes: its mode is a constructor argument, hence one class per mode), RandPermPair
// (actions/RandPermPair.cpp:67-196) and MoveStats (actions/MoveStats.cpp:107-285).  None of the three is configured by a shipped
// parameter file (MoveStats sits in RabbitPop, NewMoveStatTestPop and OoANavSHybYchMTDPop); the reference's own templates are
// added to tut_EnvironAltPop, the <prio> entries of the parameter file decide which of WeightedMove / CondWeightedMove and
// RandomPair / RandPermPair run.
template <int MODE>
class ExtProbePop : public tut_EnvironAltPop {
public:
    ExtProbePop(SCellGrid *pCG, PopFinder *pPF, int iLayerSize, IDGen **apIDG, uint32_t *aulState, uint *aiSeeds)
        : tut_EnvironAltPop(pCG, pPF, iLayerSize, apIDG, aulState, aiSeeds) {
        m_pCond = new SimpleCondition(m_pCG->m_pGeography->m_adAltitude, MODE);
        m_pCWM = new CondWeightedMove<tut_EnvironAltAgent>(this, m_pCG, "", m_apWELL, m_adEnvWeights, m_pCond);
        m_prio.addAction(m_pCWM);
        m_pRPP = new RandPermPair<tut_EnvironAltAgent>(this, m_pCG, "", m_apWELL);
        m_prio.addAction(m_pRPP);
        m_pMS = new MoveStats<tut_EnvironAltAgent>(this, m_pCG, "");
        m_prio.addAction(m_pMS);
    }
    virtual ~ExtProbePop() { delete m_pCWM; delete m_pRPP; delete m_pMS; delete m_pCond; }
    SimpleCondition *m_pCond;
    CondWeightedMove<tut_EnvironAltAgent> *m_pCWM;
    RandPermPair<tut_EnvironAltAgent> *m_pRPP;
    MoveStats<tut_EnvironAltAgent> *m_pMS;
};

// Probe class for pinning the Genetics action itself (actions/Genetics.cpp:285-337 makeOffspring: strand choice, crossover /
// free recombination of both parents, mutation count and positions, all from the action's OWN generators seeded from
// aiSeeds[1], :91-123) with 1-bit (genes/BitGeneUtils.cpp) and 2-bit (genes/GeneUtils.cpp) nucleotides.  The shipped classes
// that carry it (OoANavGenPop, OoANavGen2bitPop ...) need Climate/Vegetation/Navigation files; here the reference's own
// Genetics<T,U> is added to the reference's tut_EnvironAltPop and called from makePopSpecificOffspring exactly as
// populations/OoANavGenPop.cpp:231-245 does.  Genetics::init() is what extractAttributesQDF calls after the attributes are
// in (actions/Genetics.cpp:488-491); the XML path never calls it, the driver does.
template <class U>
class GenProbePop : public tut_EnvironAltPop {
public:
    GenProbePop(SCellGrid *pCG, PopFinder *pPF, int iLayerSize, IDGen **apIDG, uint32_t *aulState, uint *aiSeeds)
        : tut_EnvironAltPop(pCG, pPF, iLayerSize, apIDG, aulState, aiSeeds) {
        m_pGenetics = new Genetics<tut_EnvironAltAgent, U>(this, m_pCG, "", m_pAgentController, &m_vMergedDeadList, m_aiSeeds[1]);
        m_prio.addAction(m_pGenetics);
    }
    virtual ~GenProbePop() { delete m_pGenetics; }
    int makePopSpecificOffspring(int iAgent, int iMother, int iFather) {
        m_pGenetics->initialize(m_fCurTime);
        int iResult = m_pGenetics->makeOffspring(iAgent, iMother, iFather);
        tut_EnvironAltPop::makePopSpecificOffspring(iAgent, iMother, iFather);
        return iResult;
    }
    Genetics<tut_EnvironAltAgent, U> *m_pGenetics;
};

// Probe classes for pinning the other five combine modes of MultiEvaluator (actions/MultiEvaluator.cpp:263-478) and findBlockings
// (:579-598).  Shipped classes use MODE_ADD_SIMPLE and MODE_MUL_SIMPLE only (OoANavPop and six relatives, which need Genetics-free
// but QDF-bound set-ups); here the reference's own tut_EnvironCapAltPop gets its MultiEvaluator replaced by one of the given mode
// over NON-cumulating SingleEvaluators, registered as an observer -- exactly how populations/OoANavPop.cpp:50-62 builds its
// multiplicative evaluator.  MODE_MAX_BLOCK dereferences m_acAllowed, which the reference only allocates for MODE_ADD_BLOCK
// (actions/MultiEvaluator.cpp:45-47): the probe allocates the array so that the reference's code can run at all.
template <int MODE>
class MultiProbePop : public tut_EnvironCapAltPop {
public:
    typedef tut_EnvironCapAltAgent A;
    MultiProbePop(SCellGrid *pCG, PopFinder *pPF, int iLayerSize, IDGen **apIDG, uint32_t *aulState, uint *aiSeeds)
        : tut_EnvironCapAltPop(pCG, pPF, iLayerSize, apIDG, aulState, aiSeeds) {
        const std::string name = m_pME->getActionName();
        delete m_pME;
        MultiEvaluator<A>::evaluatorinfos info;
        SingleEvaluator<A> *pAlt = new SingleEvaluator<A>(this, m_pCG, "Alt", NULL, (double *)m_pGeography->m_adAltitude, "AltPref", false, EVENT_ID_GEO);
        info.push_back(std::pair<std::string, Evaluator<A> *>("Multi_weight_alt", pAlt));
        SingleEvaluator<A> *pCap = new SingleEvaluator<A>(this, m_pCG, "NPP", NULL, m_adCapacities, "", false, EVENT_ID_VEG);
        info.push_back(std::pair<std::string, Evaluator<A> *>("Multi_weight_npp", pCap));
        m_pME = new MultiEvaluator<A>(this, m_pCG, "NPP+Alt", m_adEnvWeights, info, MODE, true);
        if (MODE == MODE_MAX_BLOCK) m_pME->m_acAllowed = new uchar[m_pCG->m_iNumCells * (m_pCG->m_iConnectivity + 1)];
        addObserver(m_pME);
        m_prio.m_names[name] = m_pME;  // takes the place of the evaluator the base class registered
    }
};

struct RefSim {
    int nCells = 0;
    int nThreads = 1;
    bool quiet = true;
    SCellGrid *cg = nullptr;
    Geography *geo = nullptr;
    Climate *cli = nullptr;
    Vegetation *veg = nullptr;
    PopLooper *looper = nullptr;
    IDGen **idg = nullptr;
    PopAccess *pa = nullptr;
    bool adapter = false;  // the population keeps its agents on the GPU: preWrite brings them back before they are read
    uint32_t state[16];
    uint seeds[8];
};

}  // namespace

extern "C" {

// ---- stand-alone known-answer helpers (utils/WELL512.cpp:70-86, utils/PolyLine.cpp:60-89) ----
int qref_well_sequence(const uint32_t *state16, int n, uint32_t *out) {
    uint32_t tmp[16];
    memcpy(tmp, state16, sizeof(tmp));
    WELL512 w(tmp);
    for (int i = 0; i < n; i++) out[i] = w.wrand();
    return 0;
}

int qref_polyline_eval(const char *def, int n, const double *x, double *out, int float_cast) {
    PolyLine *pl = PolyLine::readFromString(def);
    if (pl == NULL) return -1;
    for (int i = 0; i < n; i++) out[i] = float_cast ? pl->getVal((float)x[i]) : pl->getVal(x[i]);
    delete pl;
    return 0;
}

// ---- genome primitives (genes/BitGeneUtils.cpp:57-75,116-186,190-220; utils/BinomialDist.cpp:61-103) ----
int qref_bit_crossover(const uint32_t *state16, const uint64_t *in, int genome_size, int n_cross, uint64_t *out) {
    uint32_t tmp[16]; memcpy(tmp, state16, sizeof(tmp));
    WELL512 w(tmp);
    BitGeneUtils::crossOver((ulong *)out, (const ulong *)in, genome_size, n_cross, &w);
    return 0;
}
int qref_bit_freereco(const uint32_t *state16, const uint64_t *in, int n_blocks, uint64_t *out) {
    uint32_t tmp[16]; memcpy(tmp, state16, sizeof(tmp));
    WELL512 w(tmp);
    BitGeneUtils::freeReco((ulong *)out, (ulong *)in, n_blocks, &w);
    return 0;
}
int qref_bit_mutate(const uint32_t *state16, uint64_t *genome, int n_bits, int n_mut) {
    uint32_t tmp[16]; memcpy(tmp, state16, sizeof(tmp));
    WELL512 w(tmp);
    BitGeneUtils::mutateNucs((ulong *)genome, n_bits, n_mut, &w);
    return 0;
}
int qref_binomial_table(double prob, int n, double eps, int cap, double *out) {
    BinomialDist *b = BinomialDist::create(prob, n, eps);
    if (b == NULL) return -1;
    int nb = (int)b->m_iNumBins;
    for (int i = 0; i < nb && i < cap; i++) out[i] = b->m_adLookUp[i];
    delete b;
    return nb;
}
int qref_binomial_get_n(double prob, int n, double eps, double r) {
    BinomialDist *b = BinomialDist::create(prob, n, eps);
    if (b == NULL) return -2;
    int k = b->getN(r);
    delete b;
    return k;
}

// ---- simulation ----
void *qref_create(const char *xml_path, const char *class_name, int nCells, const int *nbr6,
                  const double *altitude, const uint8_t *ice, int nThreads, const uint32_t *state16,
                  int layerSize, int quiet) {
    RefSim *s = new RefSim;
    s->nCells = nCells;
    s->nThreads = nThreads;
    s->quiet = quiet != 0;
    Quiet q(s->quiet);
    omp_set_num_threads(nThreads);

    stringmap sm;
    s->cg = new SCellGrid(0, (uint)nCells, sm);
    s->cg->m_aCells = new SCell[nCells];
    for (int c = 0; c < nCells; c++) {
        SCell &sc = s->cg->m_aCells[c];
        sc.m_iGlobalID = c;
        int nn = 0;
        for (int k = 0; k < MAX_NEIGH; k++) {
            sc.m_aNeighbors[k] = nbr6[c * MAX_NEIGH + k];
            if (sc.m_aNeighbors[k] >= 0) nn++;
        }
        sc.m_iNumNeighbors = (uchar)nn;
        s->cg->m_mIDIndexes[c] = c;
    }
    s->geo = new Geography(s->cg, (uint)nCells, 6, 6371.3);
    s->cg->setGeography(s->geo);
    for (int c = 0; c < nCells; c++) {
        s->geo->m_adAltitude[c] = altitude[c];
        s->geo->m_abIce[c] = ice ? (ice[c] != 0) : false;
    }

    memcpy(s->state, state16, sizeof(s->state));
    for (int i = 0; i < 8; i++) s->seeds[i] = 0;
    s->idg = new IDGen *[nThreads];
    for (int t = 0; t < nThreads; t++) s->idg[t] = new IDGen(0, t, nThreads);
    s->looper = new PopLooper();
    s->looper->dTimeActions = 0;
    s->looper->dTimeFinalize = 0;

    // climate and vegetation groups (needed by NPPCapacity, actions/NPPCapacity.cpp:92-115)
    s->cli = new Climate(s->cg, (uint)nCells, 1);
    s->cg->setClimate(s->cli);
    s->veg = new Vegetation(s->cg, (uint)nCells, 1);
    s->cg->setVegetation(s->veg);
    for (int c = 0; c < nCells; c++) { s->cli->m_adAnnualMeanTemp[c] = 0; s->cli->m_adAnnualRainfall[c] = 0; s->geo->m_adWater[c] = 0; }

    s->cg->setNavigation(new Navigation(s->cg));  // filled by qref_set_navigation; Navigate keeps the pointer (actions/Navigate.cpp:40)
    const int ls = layerSize > 0 ? layerSize : 65536;
    if (std::string(class_name) == "tut_EnvironAltPop") {
        s->pa = new PopAccessT<tut_EnvironAltPop, tut_EnvironAltAgent>(new tut_EnvironAltPop(s->cg, s->looper, ls, s->idg, s->state, s->seeds));
    } else if (std::string(class_name) == "tut_EnvironAltNavPop") {
        s->pa = new PopAccessT<NavProbePop, tut_EnvironAltAgent>(new NavProbePop(s->cg, s->looper, ls, s->idg, s->state, s->seeds));
    } else if (std::string(class_name) == "tut_SexualPop") {
        s->pa = new PopAccessT<tut_SexualPop, tut_SexualAgent>(new tut_SexualPop(s->cg, s->looper, ls, s->idg, s->state, s->seeds));
    } else if (std::string(class_name) == "tut_MovePop") {
        s->pa = new PopAccessT<tut_MovePop, tut_MoveAgent>(new tut_MovePop(s->cg, s->looper, ls, s->idg, s->state, s->seeds));
    } else if (std::string(class_name) == "tut_OldAgeDiePop") {
        s->pa = new PopAccessT<tut_OldAgeDiePop, tut_OldAgeDieAgent>(new tut_OldAgeDiePop(s->cg, s->looper, ls, s->idg, s->state, s->seeds));
    } else if (std::string(class_name) == "tut_EnvironAltConfPop") {
        s->pa = new PopAccessT<ConfProbePop, tut_EnvironAltAgent>(new ConfProbePop(s->cg, s->looper, ls, s->idg, s->state, s->seeds));
#define QHG_EXT_PROBE(M)                                                                                                              \
    } else if (std::string(class_name) == "tut_EnvironAltCond" #M "Pop") {                                                            \
        s->pa = new PopAccessT<ExtProbePop<M>, tut_EnvironAltAgent>(new ExtProbePop<M>(s->cg, s->looper, ls, s->idg, s->state, s->seeds));
    QHG_EXT_PROBE(0) QHG_EXT_PROBE(1) QHG_EXT_PROBE(2) QHG_EXT_PROBE(3) QHG_EXT_PROBE(4) QHG_EXT_PROBE(5) QHG_EXT_PROBE(6) QHG_EXT_PROBE(7)
#undef QHG_EXT_PROBE
    } else if (std::string(class_name) == "tut_EnvironAltVarPop") {
        s->pa = new PopAccessT<VarProbePop, tut_EnvironAltAgent>(new VarProbePop(s->cg, s->looper, ls, s->idg, s->state, s->seeds));
    } else if (std::string(class_name) == "tut_EnvironAltGenPop") {
        s->pa = new PopAccessT<GenProbePop<BitGeneUtils>, tut_EnvironAltAgent>(new GenProbePop<BitGeneUtils>(s->cg, s->looper, ls, s->idg, s->state, s->seeds));
    } else if (std::string(class_name) == "tut_EnvironAltGen2bitPop") {
        s->pa = new PopAccessT<GenProbePop<GeneUtils>, tut_EnvironAltAgent>(new GenProbePop<GeneUtils>(s->cg, s->looper, ls, s->idg, s->state, s->seeds));
    } else if (std::string(class_name) == "tut_ParthenoPop") {
        s->pa = new PopAccessT<tut_ParthenoPop, tut_ParthenoAgent>(new tut_ParthenoPop(s->cg, s->looper, ls, s->idg, s->state, s->seeds));
    } else if (std::string(class_name) == "tut_StaticPop") {
        s->pa = new PopAccessT<tut_StaticPop, Agent>(new tut_StaticPop(s->cg, s->looper, ls, s->idg, s->state, s->seeds));
    } else if (std::string(class_name) == "tut_EnvironCapAltAddBlockPop") {
        s->pa = new PopAccessT<MultiProbePop<MODE_ADD_BLOCK>, tut_EnvironCapAltAgent>(new MultiProbePop<MODE_ADD_BLOCK>(s->cg, s->looper, ls, s->idg, s->state, s->seeds));
    } else if (std::string(class_name) == "tut_EnvironCapAltMulPop") {
        s->pa = new PopAccessT<MultiProbePop<MODE_MUL_SIMPLE>, tut_EnvironCapAltAgent>(new MultiProbePop<MODE_MUL_SIMPLE>(s->cg, s->looper, ls, s->idg, s->state, s->seeds));
    } else if (std::string(class_name) == "tut_EnvironCapAltMaxPop") {
        s->pa = new PopAccessT<MultiProbePop<MODE_MAX_SIMPLE>, tut_EnvironCapAltAgent>(new MultiProbePop<MODE_MAX_SIMPLE>(s->cg, s->looper, ls, s->idg, s->state, s->seeds));
    } else if (std::string(class_name) == "tut_EnvironCapAltMaxBlockPop") {
        s->pa = new PopAccessT<MultiProbePop<MODE_MAX_BLOCK>, tut_EnvironCapAltAgent>(new MultiProbePop<MODE_MAX_BLOCK>(s->cg, s->looper, ls, s->idg, s->state, s->seeds));
    } else if (std::string(class_name) == "tut_EnvironCapAltMinPop") {
        s->pa = new PopAccessT<MultiProbePop<MODE_MIN_SIMPLE>, tut_EnvironCapAltAgent>(new MultiProbePop<MODE_MIN_SIMPLE>(s->cg, s->looper, ls, s->idg, s->state, s->seeds));
    } else if (std::string(class_name) == "OoANavGenPop") {  // populations/OoANavGenPop.cpp:33-97, the shipped class itself
        s->pa = new PopAccessT<OoANavGenPop, OoANavGenAgent>(new OoANavGenPop(s->cg, s->looper, ls, s->idg, s->state, s->seeds));
    } else if (std::string(class_name) == "tut_EnvironCapAltPop") {
        s->pa = new PopAccessT<tut_EnvironCapAltPop, tut_EnvironCapAltAgent>(new tut_EnvironCapAltPop(s->cg, s->looper, ls, s->idg, s->state, s->seeds));
#ifdef QHG_WITH_GPU_ADAPTER
    } else if (std::string(class_name) == "tut_EnvironAltGpuPop") {  // the reference's loop drives the CUDA path
        s->pa = new PopAccessT<tut_EnvironAltGpuPop, tut_EnvironAltAgent>(new tut_EnvironAltGpuPop(s->cg, s->looper, ls, s->idg, s->state, s->seeds));
        s->adapter = true;
        class_name = "tut_EnvironAltPop";  // the class entry of the parameter file
    } else if (std::string(class_name) == "tut_EnvironCapAltGpuPop") {  // integration/qhg_gpu_pop.h over tut_EnvironCapAltPop
        s->pa = new PopAccessT<tut_EnvironCapAltGpuPop, tut_EnvironCapAltAgent>(new tut_EnvironCapAltGpuPop(s->cg, s->looper, ls, s->idg, s->state, s->seeds));
        s->adapter = true;
        class_name = "tut_EnvironCapAltPop";
    } else if (std::string(class_name) == "OoANavGenGpuPop") {          // ... over OoANavGenPop
        s->pa = new PopAccessT<OoANavGenGpuPop, OoANavGenAgent>(new OoANavGenGpuPop(s->cg, s->looper, ls, s->idg, s->state, s->seeds));
        s->adapter = true;
        class_name = "OoANavGenPop";
    } else if (std::string(class_name) == "dyn:tut_EnvironAltGpuPop") {
        // the way the application gets a plugin population (app/SimParams.cpp:1395-1407 with --dyn-pops --so-dirs): the
        // reference's DynPopFactory scans the directory for *Wrapper.so, dlopens it and calls its createPop
        // (populations/DynPopFactory.cpp:80-163).  The plugin is integration/tut_EnvironAltGpuPopWrapper.cpp.
        const char *dir = getenv("QHG_REF_SO_DIR");
        stringvec vDirs;
        vDirs.push_back(dir ? dir : ".");
        DynPopFactory *pF = DynPopFactory::createInstance(vDirs, s->cg, s->looper, ls, s->idg, s->state, s->seeds);
        PopBase *pPB = (pF != NULL) ? pF->createPopulationByName("tut_EnvironAltGpuPop") : NULL;
        if (pPB == NULL) { fprintf(stderr, "[qref_create] DynPopFactory could not create [tut_EnvironAltGpuPop] from [%s]\n", dir ? dir : "."); return NULL; }
        s->pa = new PopAccessT<tut_EnvironAltGpuPop, tut_EnvironAltAgent>(static_cast<tut_EnvironAltGpuPop *>(pPB));
        s->adapter = true;
        class_name = "tut_EnvironAltPop";  // (the factory stays alive: it keeps the library handle, populations/DynPopFactory.cpp:38-42)
#endif
    } else {
        fprintf(stderr, "[qref_create] unknown population class [%s]\n", class_name);
        return NULL;
    }
    ParamProvider2 *pp = ParamProvider2::createInstance(xml_path);
    int rc = -1;
    if (pp != NULL) {
        rc = pp->selectClass(class_name);
        if (rc == 0) rc = s->pa->base()->readSpeciesData(pp);
        delete pp;
    }
    if (rc != 0) {
        fprintf(stderr, "[qref_create] could not read species data from [%s] class [%s]\n", xml_path, class_name);
        return NULL;
    }
    return s;
}

int qref_add_agents(void *h, long n, const int *cell, const int64_t *id, const float *birth,
                    const uint8_t *gender, const float *age, const float *lastBirth, const uint32_t *life) {
    RefSim *s = (RefSim *)h;
    Quiet q(s->quiet);
    long done = 0;
    while (done < n) {  // reserveAgentSpace takes int
        int chunk = (int)((n - done > (1 << 30)) ? (1 << 30) : (n - done));
        int start = s->pa->reserve(chunk);
#pragma omp parallel for
        for (int i = 0; i < chunk; i++) {
            long j = done + i;
            AgentRec r;
            r.life = life ? life[j] : LIFE_STATE_ALIVE; r.cell = cell[j]; r.id = id[j]; r.birth = birth[j];
            r.gender = gender[j]; r.age = age[j]; r.lastBirth = lastBirth[j];
            s->pa->put(start + i, r, s->cg->m_aCells[cell[j]].m_iGlobalID);
        }
        for (int i = 0; i < chunk; i++) {
            if (id[done + i] > s->pa->maxID()) s->pa->maxID() = id[done + i];
        }
        done += chunk;
    }
    return 0;
}

// the Navigation group (core/Navigation.h:22-39) as io/NavGroupReader would fill it: ports with their destination cells
// and distances (CSR), manual bridges; call before qref_start
int qref_set_navigation(void *h, int nPorts, const int *portCell, const int *portPtr, const int *destCell, const double *dist,
                        int nBridges, const int *bridges) {
    RefSim *s = (RefSim *)h;
    distancemap dm;
    for (int i = 0; i < nPorts; i++) {
        distlist dl;
        for (int k = portPtr[i]; k < portPtr[i + 1]; k++) dl[destCell[k]] = dist[k];
        dm[portCell[i]] = dl;
    }
    bridgelist bl;
    for (int b = 0; b < nBridges; b++) bl.push_back(bridgedef(bridges[2 * b], bridges[2 * b + 1]));
    s->cg->m_pNavigation->setData(dm, 1.0);
    s->cg->m_pNavigation->setBridges(bl);
    return 0;
}

int qref_start(void *h) {
    RefSim *s = (RefSim *)h;
    Quiet q(s->quiet);
    int rc = s->looper->addPop(s->pa->base());
    idtype maxID = s->looper->getMaxID();
    for (int t = 0; t < s->nThreads; t++) s->idg[t]->setData(maxID + 1, t, s->nThreads);
    rc += s->looper->preLoop();
    return rc;
}

int qref_step(void *h, float t) {
    RefSim *s = (RefSim *)h;
    Quiet q(s->quiet);
    return s->looper->doStep(t);
}

// n steps back to back, t = t0, t0+1, ...; returns wall seconds of the doStep calls only
double qref_run(void *h, float t0, int nSteps, int64_t *agentSteps) {
    RefSim *s = (RefSim *)h;
    Quiet q(s->quiet);
    int64_t as = 0;
    double w0 = omp_get_wtime();
    for (int i = 0; i < nSteps; i++) {
        as += (int64_t)s->pa->base()->getNumAgentsEffective();
        s->looper->doStep(t0 + i);
    }
    double w1 = omp_get_wtime();
    if (agentSteps) *agentSteps = as;
    return w1 - w0;
}

long qref_num_agents(void *h) { return (long)((RefSim *)h)->pa->base()->getNumAgentsEffective(); }

// live agents in slot order; returns number written (or needed if cap too small)
long qref_get_agents(void *h, long cap, int *cell, int64_t *id, float *birth, uint8_t *gender, float *age,
                     float *lastBirth, uint32_t *life, int *mate, int *slot) {
    RefSim *s = (RefSim *)h;
    if (s->adapter) { Quiet q(s->quiet); s->pa->base()->preWrite(0.0f); }
    int first = s->pa->first();
    if (first < 0) return 0;
    int last = s->pa->last();
    long k = 0;
    for (int i = first; i <= last; i++) {
        AgentRec a;
        if (!s->pa->get(i, a)) continue;
        if (k < cap) {
            if (cell) cell[k] = a.cell;
            if (id) id[k] = a.id;
            if (birth) birth[k] = a.birth;
            if (gender) gender[k] = a.gender;
            if (age) age[k] = a.age;
            if (lastBirth) lastBirth[k] = a.lastBirth;
            if (life) life[k] = a.life;
            if (mate) mate[k] = a.mate;
            if (slot) slot[k] = i;
        }
        k++;
    }
    return k;
}

// ---- genomes of the Genetics probe populations ----
int qref_genetics_init(void *h, int genomeSize, int numCrossOvers, double mutationRate) {
    RefSim *s = (RefSim *)h;
    Quiet q(s->quiet);
    return s->pa->geneticsInit(genomeSize, numCrossOvers, mutationRate);
}
// rows for the agents in slots [firstSlot, firstSlot + n): what readAdditionalDataQDF would load beside the agents
int qref_set_genomes(void *h, int firstSlot, long n, const uint64_t *rows) {
    RefSim *s = (RefSim *)h;
    const int w = s->pa->genomeWords();
    if (w <= 0) return -1;
    for (long i = 0; i < n; i++) memcpy(s->pa->genomeRow(firstSlot + (int)i), rows + (size_t)i * w, sizeof(uint64_t) * w);
    return w;
}
// rows of the live agents in slot order (the order of qref_get_agents); returns the words per row
long qref_get_genomes(void *h, long cap, uint64_t *rows) {
    RefSim *s = (RefSim *)h;
    if (s->adapter) { Quiet q(s->quiet); s->pa->base()->preWrite(0.0f); }
    const int w = s->pa->genomeWords();
    if (w <= 0) return -1;
    int first = s->pa->first();
    if (first < 0) return w;
    long k = 0;
    for (int i = first; i <= s->pa->last(); i++) {
        AgentRec a;
        if (!s->pa->get(i, a)) continue;
        if (k < cap) memcpy(rows + (size_t)k * w, s->pa->genomeRow(i), sizeof(uint64_t) * w);
        k++;
    }
    return w;
}
// m_iNumBabies of every live agent, in the order of qref_get_agents (populations/OoANavGenPop.cpp:243); -1: the class has none
long qref_get_num_babies(void *h, long cap, int *out) {
    RefSim *s = (RefSim *)h;
    if (s->adapter) { Quiet q(s->quiet); s->pa->base()->preWrite(0.0f); }
    int first = s->pa->first();
    if (first < 0) return 0;
    long k = 0;
    for (int i = first; i <= s->pa->last(); i++) {
        AgentRec a;
        if (!s->pa->get(i, a)) continue;
        if (k < cap) out[k] = s->pa->numBabies(i);
        k++;
    }
    return k;
}
// is this population class part of this build?
int qref_has_class(const char *name) {
    static const char *const known[] = {"tut_EnvironAltPop", "tut_EnvironAltNavPop", "tut_SexualPop", "tut_MovePop", "tut_OldAgeDiePop",
                                        "tut_EnvironAltConfPop", "tut_EnvironAltVarPop", "tut_EnvironAltGenPop", "tut_EnvironAltGen2bitPop",
                                        "tut_ParthenoPop", "tut_StaticPop", "tut_EnvironCapAltPop", "OoANavGenPop", "tut_EnvironCapAltAddBlockPop",
                                        "tut_EnvironCapAltMulPop", "tut_EnvironCapAltMaxPop", "tut_EnvironCapAltMaxBlockPop", "tut_EnvironCapAltMinPop"};
    for (const char *k : known) if (std::string(k) == name) return 1;
    const std::string n(name);
    if (n.size() == 22 && n.rfind("tut_EnvironAltCond", 0) == 0 && n.substr(19) == "Pop" && n[18] >= '0' && n[18] <= '7') return 1;
    return 0;
}
// PopBase::modifyAttributes(name, value) (core/SPopulation.cpp "modifyAttributes": forwarded to every action)
int qref_modify_attribute(void *h, const char *name, double value) {
    RefSim *s = (RefSim *)h;
    Quiet q(s->quiet);
    return s->pa->base()->modifyAttributes(name, value);
}
// state of the Genetics action's generator of thread 0 (built from aiSeeds[1] by WELLUtils::buildWELLs, MD5 of seed phrases)
int qref_genetics_well(void *h, uint32_t *state16, uint32_t *index) {
    RefSim *s = (RefSim *)h;
    WELL512 *w = s->pa->geneticsWell();
    if (w == NULL) return -1;
    memcpy(state16, w->getState(), sizeof(uint32_t) * 16);
    *index = w->getIndex();
    return 0;
}

// ---- stand-alone 2-bit genome primitives (genes/GeneUtils.cpp:112-146,184-245,322-362) ----
int qref_gene2_crossover(const uint32_t *state16, const uint64_t *in, int genome_size, int n_cross, uint64_t *out) {
    uint32_t tmp[16]; memcpy(tmp, state16, sizeof(tmp));
    WELL512 w(tmp);
    GeneUtils::crossOver((ulong *)out, (const ulong *)in, genome_size, n_cross, &w);
    return 0;
}
int qref_gene2_freereco(const uint32_t *state16, const uint64_t *in, int n_blocks, uint64_t *out) {
    uint32_t tmp[16]; memcpy(tmp, state16, sizeof(tmp));
    WELL512 w(tmp);
    GeneUtils::freeReco((ulong *)out, (ulong *)in, n_blocks, &w);
    return 0;
}
int qref_gene2_mutate(const uint32_t *state16, uint64_t *genome, int n_nucs, int n_mut) {
    uint32_t tmp[16]; memcpy(tmp, state16, sizeof(tmp));
    WELL512 w(tmp);
    GeneUtils::mutateNucs((ulong *)genome, n_nucs, n_mut, &w);
    return 0;
}

int qref_get_counts(void *h, uint64_t *out) {
    RefSim *s = (RefSim *)h;
    if (s->adapter) s->pa->base()->updateNumAgentsPerCell();
    for (int c = 0; c < s->nCells; c++) out[c] = s->adapter ? s->pa->cellCounts()[c] : s->pa->base()->getNumAgents(c);
    return 0;
}

int qref_get_weights(void *h, double *out) {  // nCells*7, actions/SingleEvaluator.cpp:174-243
    RefSim *s = (RefSim *)h;
    memcpy(out, s->pa->envWeights(), sizeof(double) * (size_t)s->nCells * 7);
    return 0;
}

int qref_get_bd(void *h, double *b, double *d) {  // actions/LinearBirth.cpp:97-112, LinearDeath.cpp:101-119
    RefSim *s = (RefSim *)h;
    double *pb = NULL, *pd = NULL;
    if (s->pa->bd(pb, pd) != 0) return -1;
    memcpy(b, pb, sizeof(double) * s->nCells);
    memcpy(d, pd, sizeof(double) * s->nCells);
    return 0;
}

// carrying capacities of NPPCapacity (actions/NPPCapacity.cpp:138-217); NULL array for populations without one
int qref_get_capacities(void *h, double *out) {
    RefSim *s = (RefSim *)h;
    double *k = s->pa->capacities();
    if (k == NULL) return -1;
    memcpy(out, k, sizeof(double) * s->nCells);
    return 0;
}

// ---- the agent dataset of a QDF file, through the reference's own writers and readers ---------------------------------
// HDF5 is not installed here; oracle/stubs/hdf5_stubs.cpp backs the handful of calls of this path with memory.
extern "C" long qhgstub_dataset_info(hid_t dset, size_t *elem, hid_t *type);
extern "C" int qhgstub_dataset_bytes(hid_t dset, void *out);
extern "C" hid_t qhgstub_dataset_from_bytes(hid_t type, hsize_t n, const void *bytes);
extern "C" int qhgstub_type_member(hid_t type, int i, char *name, int cap, size_t *offset, hid_t *mtype);

// what PopLooper::preWrite (core/PopLooper.cpp:146-157, called from app/Simulator.cpp:601) and the agent part of PopWriter::write
// (io/PopWriter.cpp:84-118) do for one population: preWrite, a dataspace of getNumAgentsEffective() records, the dataset,
// PopBase::writeAgentDataQDF.  Returns the dataset handle (< 0: failure).
long long qref_qdf_write_agents(void *h, float t) {
    RefSim *s = (RefSim *)h;
    Quiet q(s->quiet);
    PopBase *pb = s->pa->base();
    if (pb->preWrite(t) != 0) return -2;
    hsize_t dims = pb->getNumAgentsEffective();
    hid_t hSpace = H5Screate_simple(1, &dims, NULL);
    hid_t hType = pb->getAgentQDFDataType();
    hid_t hSet = H5Dcreate2(0, "AgentDataSet", hType, hSpace, H5P_DEFAULT, H5P_DEFAULT, H5P_DEFAULT);
    if (hSet < 0) return -3;
    const int rc = pb->writeAgentDataQDF(hSpace, hSet, hType);
    H5Sclose(hSpace);
    return rc == 0 ? (long long)hSet : -4;
}
// records and record size of a dataset; its bytes if `out` is not NULL
long qref_qdf_dataset(long long dset, size_t *elem, void *out) {
    const long n = qhgstub_dataset_info((hid_t)dset, elem, NULL);
    if (n >= 0 && out != NULL) qhgstub_dataset_bytes((hid_t)dset, out);
    return n;
}
// member i of the compound type the population registered for its agents (core/SPopulation.cpp:1356-1372 + the class's
// addPopSpecificAgentDataTypeQDF): name, offset, stub type code (oracle/stubs/hdf5.h)
int qref_qdf_agent_member(void *h, int i, char *name, int cap, size_t *offset, int *typeCode) {
    RefSim *s = (RefSim *)h;
    hid_t hType = s->pa->base()->getAgentQDFDataType();
    if (hType <= 0) hType = s->pa->base()->createAgentDataTypeQDF();
    hid_t m = -1;
    const int rc = qhgstub_type_member(hType, i, name, cap, offset, &m);
    *typeCode = (int)m;
    return rc;
}
// the agent part of PopReader::read (io/PopReader.cpp:143-170) on a dataset holding `n` records: PopBase::readAgentDataQDF
int qref_qdf_read_agents(void *h, long n, const void *bytes) {
    RefSim *s = (RefSim *)h;
    Quiet q(s->quiet);
    PopBase *pb = s->pa->base();
    hid_t hType = pb->getAgentQDFDataType();
    if (hType <= 0) hType = pb->createAgentDataTypeQDF();
    hid_t hSet = qhgstub_dataset_from_bytes(hType, (hsize_t)n, bytes);
    if (hSet < 0) return -2;
    hid_t hSpace = H5Dget_space(hSet);
    const int rc = pb->readAgentDataQDF(hSpace, hSet, hType);
    H5Sclose(hSpace);
    return rc;
}

// MoveStats' per-cell arrays (actions/MoveStats.h:49-51); -1 for populations without the action or before preLoop
int qref_get_move_stats(void *h, int *hops, double *dist, double *time) {
    RefSim *s = (RefSim *)h;
    int *ph; double *pd, *pt;
    if (s->pa->moveStats(ph, pd, pt) != 0) return -1;
    memcpy(hops, ph, sizeof(int) * s->nCells);
    memcpy(dist, pd, sizeof(double) * s->nCells);
    memcpy(time, pt, sizeof(double) * s->nCells);
    return 0;
}

// any per-cell environment array the supported populations read (core/Geography.h:31-39, core/Climate.h, core/Vegetation.h)
int qref_set_env(void *h, const char *name, const double *v) {
    RefSim *s = (RefSim *)h;
    std::string n(name);
    for (int c = 0; c < s->nCells; c++) {
        if (n == "Altitude") s->geo->m_adAltitude[c] = v[c];
        else if (n == "Ice") s->geo->m_abIce[c] = v[c] != 0;
        else if (n == "Water") s->geo->m_adWater[c] = v[c];
        else if (n == "Coastal") s->geo->m_abCoastal[c] = v[c] != 0;
        else if (n == "Latitude") s->geo->m_adLatitude[c] = v[c];
        else if (n == "Longitude") s->geo->m_adLongitude[c] = v[c];
        else if (n == "AnnualMeanTemp") s->cli->m_adAnnualMeanTemp[c] = v[c];
        else if (n == "AnnualRainfall") s->cli->m_adAnnualRainfall[c] = v[c];
        else if (n == "BaseNPP") s->veg->m_adBaseANPP[c] = v[c];
        else return -1;
    }
    return 0;
}

// deliver an event the way app/Simulator.cpp:728-735,372-374 does: updateEvent for the id(s), then flushEvents
int qref_event(void *h, int event_id, float t, int flush) {
    RefSim *s = (RefSim *)h;
    Quiet q(s->quiet);
    int rc = s->pa->base()->updateEvent(event_id, NULL, t);
    if (flush) s->pa->base()->flushEvents(t);
    return rc;
}

// ATanDeath probability exactly as actions/ATanDeath.cpp:49-59,75 computes it (fAge is the agent's float age)
int qref_atan_prob(void *h, int n, const float *age, double *p) {
    RefSim *s = (RefSim *)h;
    double scale, slope, maxAge;
    s->pa->atanParams(scale, slope, maxAge);
    for (int i = 0; i < n; i++) {
        p[i] = 0.5 + scale * atan(slope * (age[i] - maxAge)) / Q_PI;
    }
    return 0;
}

// change env arrays and deliver EVENT_ID_GEO like app/Simulator.cpp:728-735 does
int qref_geo_event(void *h, const double *altitude, const uint8_t *ice, float t) {
    RefSim *s = (RefSim *)h;
    Quiet q(s->quiet);
    for (int c = 0; c < s->nCells; c++) {
        if (altitude) s->geo->m_adAltitude[c] = altitude[c];
        if (ice) s->geo->m_abIce[c] = ice[c] != 0;
    }
    int rc = s->pa->base()->updateEvent(EVENT_ID_GEO, NULL, t);
    s->pa->base()->flushEvents(t);
    return rc;
}

int qref_timers(void *h, double *actions, double *finalize) {
    RefSim *s = (RefSim *)h;
    *actions = s->looper->dTimeActions;
    *finalize = s->looper->dTimeFinalize;
    return 0;
}

int qref_max_threads() { return omp_get_num_procs(); }

void qref_destroy(void *h) {
    RefSim *s = (RefSim *)h;
    if (!s) return;
    Quiet q(s->quiet);
    delete s->looper;  // deletes the pops (core/PopLooper.cpp:25-34)
    for (int t = 0; t < s->nThreads; t++) delete s->idg[t];
    delete[] s->idg;
    delete s->pa;
    s->cg->delGeography();
    s->cg->delClimate();
    s->cg->delVegetation();
    delete s->cg;
    delete s;
}

}  // extern "C"
