// integration/qhg_gpu_pop.h -- the plugin adapter as a template: any shipped population class whose action set the CUDA
// library knows (include/qhg_b200.h, qhgb_create) becomes a GPU population by deriving from it.
//
//   class tut_EnvironCapAltGpuPop : public QhgGpuPop<tut_EnvironCapAltPop, tut_EnvironCapAltAgent> { ... "tut_EnvironCapAltPop" ... };
//
// The base class keeps everything the host needs unchanged -- XML / QDF parameter handling, priorities, addAgent,
// readAgentDataQDF, writeAgentDataQDFSafe, mergePop -- and this template overrides exactly the PopBase virtuals of the step
// loop (core/PopLooper.cpp:166-202, app/Simulator.cpp:303-374,601): preLoop uploads grid, environment, parameters, agents
// (and genomes, navigation); initializeStep / doActions / finalizeStep forward to the device; updateEvent / flushEvents
// re-upload the arrays the host has just re-read; preWrite brings the device state back into m_aAgents (and the genome buffer)
// so that the unchanged QDF writers work.  integration/tut_EnvironAltGpuPop.h is the same thing written out for one class.
#ifndef __QHG_GPU_POP_H__
#define __QHG_GPU_POP_H__
#include <cstdlib>
#include <map>
#include <vector>
#include "LBController.h"
#include "Geography.h"
#include "Climate.h"
#include "Vegetation.h"
#include "Navigation.h"
#include "ParamProvider2.h"
extern "C" {
#include "qhg_b200.h"
}

// Genetics keeps m_iNumCrossOvers / m_dMutationRate protected and has no getters (actions/Genetics.h:73-106); a derived class
// may form pointers to them.  (A maintainer would rather add two one-line getters next to getGenomeSize(), :64.)
template <class G>
struct QhgGeneticsPeek : public G {
    static int numCrossOvers(G *g) { return g->*(&QhgGeneticsPeek::m_iNumCrossOvers); }
    static double mutationRate(G *g) { return g->*(&QhgGeneticsPeek::m_dMutationRate); }
};

template <class BasePop, class AgentT>
class QhgGpuPop : public BasePop {
public:
    QhgGpuPop(const char *sDeviceClass, SCellGrid *pCG, PopFinder *pPF, int iLayerSize, IDGen **apIDG, uint32_t *aulState, uint *aiSeeds)
        : BasePop(pCG, pPF, iLayerSize, apIDG, aulState, aiSeeds), m_gpu(NULL) {
        // (aiSeeds is not dereferenced here: DynPopFactory calls createPop with six arguments, populations/DynPopFactory.cpp:18,150)
        int iDev = getenv("QHG_DEVICE") ? atoi(getenv("QHG_DEVICE")) : 0;
        if (qhgb_create(sDeviceClass, iDev, pCG->m_iNumCells, pCG->m_iConnectivity, 0, &m_gpu) != 0) {
            xha_printf("[QhgGpuPop] %s\n", qhgb_last_error());   // no CPU fallback: the population is unusable
        } else {
            qhgb_set_seed(m_gpu, aulState);
        }
    }
    virtual ~QhgGpuPop() { qhgb_destroy(m_gpu); }

    // XML parameters: the base class fills the Action objects; the same strings go to the device
    virtual int readSpeciesData(ParamProvider2 *pPP) {
        int iResult = BasePop::readSpeciesData(pPP);
        const classinfo *pCI = pPP->getClassInfo();
        for (stringmap::const_iterator it = pCI->prios.begin(); iResult == 0 && it != pCI->prios.end(); ++it)
            iResult += qhgb_set_prio(m_gpu, it->first.c_str(), atoi(it->second.c_str()));
        for (modulemap::const_iterator im = pCI->mods.begin(); iResult == 0 && im != pCI->mods.end(); ++im)
            iResult += pushParams(im->second);
        return iResult;
    }
    int pushParams(const ModuleComplex *pM) {  // a module and its sub-modules (the evaluators inside a MultiEvaluator)
        int iResult = 0;
        const stringmap &mA = pM->getAttributes();
        for (stringmap::const_iterator ip = mA.begin(); ip != mA.end(); ++ip)
            iResult += qhgb_set_attribute_str(m_gpu, ip->first.c_str(), ip->second.c_str());
        const modulemap &mS = pM->getSubModules();
        for (modulemap::const_iterator is = mS.begin(); is != mS.end(); ++is) iResult += pushParams(is->second);
        return iResult;
    }

    // the per-cell arrays the class's actions read (Geography always; Climate / Vegetation when NPPCapacity is there)
    int pushEnvironment() {
        const int n = this->m_iNumCells;
        Geography *pG = this->m_pCG->m_pGeography;
        std::vector<double> v(n);
        int iResult = qhgb_set_env_array(m_gpu, "Altitude", pG->m_adAltitude, n);
        for (int c = 0; c < n; c++) v[c] = pG->m_abIce[c];
        iResult += qhgb_set_env_array(m_gpu, "Ice", v.data(), n);
        if constexpr (requires { this->m_pNPPCap; }) {
            iResult += qhgb_set_env_array(m_gpu, "Water", pG->m_adWater, n);
            for (int c = 0; c < n; c++) v[c] = pG->m_abCoastal[c];
            iResult += qhgb_set_env_array(m_gpu, "Coastal", v.data(), n);
            iResult += qhgb_set_env_array(m_gpu, "Latitude", pG->m_adLatitude, n);
            iResult += qhgb_set_env_array(m_gpu, "Longitude", pG->m_adLongitude, n);
            if (this->m_pCG->m_pClimate != NULL) {
                iResult += qhgb_set_env_array(m_gpu, "AnnualMeanTemp", this->m_pCG->m_pClimate->m_adAnnualMeanTemp, n);
                iResult += qhgb_set_env_array(m_gpu, "AnnualRainfall", this->m_pCG->m_pClimate->m_adAnnualRainfall, n);
            }
            if (this->m_pCG->m_pVegetation != NULL) iResult += qhgb_set_env_array(m_gpu, "BaseNPP", this->m_pCG->m_pVegetation->m_adBaseANPP, n);
        }
        return iResult;
    }
    // the Navigation group (core/Navigation.h:13-39) as the CSR tables of qhgb_set_navigation
    int pushNavigation() {
        if constexpr (requires { this->m_pNavigate; }) {
            Navigation *pN = this->m_pCG->m_pNavigation;
            if (pN == NULL) return 0;
            std::vector<int32_t> vPort, vPtr(1, 0), vDest, vBr;
            std::vector<double> vDist;
            for (distancemap::const_iterator ip = pN->m_mDestinations.begin(); ip != pN->m_mDestinations.end(); ++ip) {
                vPort.push_back(ip->first);
                for (distlist::const_iterator id = ip->second.begin(); id != ip->second.end(); ++id) { vDest.push_back(id->first); vDist.push_back(id->second); }
                vPtr.push_back((int32_t)vDest.size());
            }
            for (size_t b = 0; b < pN->m_vBridges.size(); b++) { vBr.push_back(pN->m_vBridges[b].first); vBr.push_back(pN->m_vBridges[b].second); }
            return qhgb_set_navigation(m_gpu, (int)vPort.size(), vPort.data(), vPtr.data(), vDest.data(), vDist.data(), (int)vBr.size() / 2, vBr.data());
        } else {
            return 0;
        }
    }

    virtual int preLoop() {
        int iResult = 0;
        if constexpr (requires { this->m_pGenetics; }) {  // the genome layout must be known before the first agent arrives
            typedef typename std::remove_pointer<decltype(this->m_pGenetics)>::type G;
            iResult += qhgb_set_attribute(m_gpu, "Genetics_genome_size", this->m_pGenetics->getGenomeSize());
            iResult += qhgb_set_attribute(m_gpu, "Genetics_num_crossover", QhgGeneticsPeek<G>::numCrossOvers(this->m_pGenetics));
            iResult += qhgb_set_attribute(m_gpu, "Genetics_mutation_rate", QhgGeneticsPeek<G>::mutationRate(this->m_pGenetics));
        }
        iResult += BasePop::preLoop();  // (creates the initial genomes if the class was told to)
        const int n = this->m_iNumCells;
        std::vector<int32_t> vN(n * 6), vID(n);
        for (int c = 0; c < n; c++) {
            vID[c] = this->m_pCG->m_aCells[c].m_iGlobalID;
            for (int k = 0; k < 6; k++) vN[6 * c + k] = this->m_pCG->m_aCells[c].m_aNeighbors[k];
        }
        iResult += qhgb_set_cells(m_gpu, vN.data(), vID.data());
        iResult += pushEnvironment();
        iResult += pushNavigation();
        // agents loaded by addAgent / readAgentDataQDF -> structure of arrays (+ their genome rows, in the same order)
        std::vector<int32_t> c; std::vector<int64_t> id; std::vector<float> b, a, l; std::vector<uint8_t> g; std::vector<uint32_t> s;
        std::vector<uint64_t> vGen;
        int i0 = this->getFirstAgentIndex();
        if (i0 != LBController::NIL) for (int i = i0; i <= this->getLastAgentIndex(); i++) {
            AgentT &ag = this->m_aAgents[i];
            if (ag.m_iLifeState == LIFE_STATE_DEAD) continue;
            c.push_back(ag.m_iCellIndex); id.push_back(ag.m_ulID); b.push_back(ag.m_fBirthTime); g.push_back(ag.m_iGender);
            a.push_back(ag.m_fAge); l.push_back(ag.m_fLastBirth); s.push_back(ag.m_iLifeState);
            if constexpr (requires { this->m_pGenetics; }) {
                const ulong *pRow = this->m_pGenetics->getGenome((uint)i);
                vGen.insert(vGen.end(), pRow, pRow + 2 * this->m_pGenetics->getNumBlocks());
            }
        }
        iResult += qhgb_add_agents(m_gpu, c.size(), c.data(), id.data(), b.data(), g.data(), a.data(), l.data(), s.data());
        if constexpr (requires { this->m_pGenetics; }) {
            if (!c.empty()) iResult += qhgb_set_genomes(m_gpu, (int64_t)c.size(), vGen.data());
        }
        iResult += qhgb_pre_loop(m_gpu);
        return iResult;
    }

    virtual int initializeStep(float fTime) { this->m_fCurTime = fTime; return qhgb_initialize_step(m_gpu, fTime); }
    virtual int doActions(uint iPrio, float fTime) { return qhgb_do_actions(m_gpu, iPrio, fTime); }
    virtual int finalizeStep() {
        int iResult = qhgb_finalize_step(m_gpu);
        qhgb_step_stats st;
        qhgb_get_step_stats(m_gpu, &st);
        this->m_iNumBirths += st.births; this->m_iNumDeaths += st.deaths; this->m_iNumMoves += st.moves;
        return iResult;
    }
    virtual ulong getNumAgentsEffective() { return qhgb_get_num_agents_effective(m_gpu); }
    virtual ulong getNumAgentsTotal()     { return qhgb_get_num_agents_effective(m_gpu); }
    virtual void  updateNumAgentsPerCell() { qhgb_get_num_agents_array(m_gpu, (uint64_t *)this->m_aiNumAgentsPerCell); }
    virtual ulong getNumAgents(int iCell)  { updateNumAgentsPerCell(); return this->m_aiNumAgentsPerCell[iCell]; }

    // the arrays were re-read by Simulator::handleEnvironmentEvent (app/Simulator.cpp:668-820) before the event is delivered
    virtual int updateEvent(int iEventID, char *pData, float fT) {
        int iResult = pushEnvironment();
        if (iEventID == EVENT_ID_NAV) iResult += pushNavigation();
        iResult += qhgb_update_event(m_gpu, iEventID, fT);
        return iResult;
    }
    virtual void flushEvents(float fT) { qhgb_flush_events(m_gpu, fT); }

    // the host calls PopLooper::preWrite right before every QDF write / dump (app/Simulator.cpp:601): bring the device state
    // back into m_aAgents (and the genome buffer) so that the unchanged writeAgentDataQDFSafe / writeAdditionalDataQDF work
    virtual int preWrite(float fTime) {
        int64_t n = qhgb_get_num_agents_effective(m_gpu);
        std::vector<int32_t> c(n), cid(n); std::vector<int64_t> id(n); std::vector<float> b(n), a(n), l(n);
        std::vector<uint8_t> g(n); std::vector<uint32_t> s(n);
        if (qhgb_get_agents(m_gpu, n, c.data(), cid.data(), id.data(), b.data(), g.data(), a.data(), l.data(), s.data(), NULL) != n) return -1;
        std::vector<uint64_t> vGen;
        std::vector<int32_t> vBabies;
        int iRow = 0;
        if constexpr (requires { this->m_pGenetics; }) {
            iRow = 2 * this->m_pGenetics->getNumBlocks();
            vGen.resize((size_t)n * iRow); vBabies.resize(n);
            if (n > 0 && qhgb_get_genomes(m_gpu, n, vGen.data(), vBabies.data()) != n) return -1;
        }
        this->m_pAgentController->clear();
        int iStart = n > 0 ? this->reserveAgentSpace((int)n) : 0;
        for (int64_t i = 0; i < n; i++) {
            AgentT &ag = this->m_aAgents[iStart + i];
            ag.m_iLifeState = s[i]; ag.m_iCellIndex = c[i]; ag.m_ulCellID = cid[i]; ag.m_ulID = id[i];
            ag.m_fBirthTime = b[i]; ag.m_iGender = g[i]; ag.m_fAge = a[i]; ag.m_fLastBirth = l[i]; ag.m_iMateIndex = -3;
            if constexpr (requires { this->m_pGenetics; }) {
                ag.m_iNumBabies = vBabies[i];
                memcpy(this->m_pGenetics->getGenome((uint)(iStart + i)), &vGen[(size_t)i * iRow], sizeof(uint64_t) * iRow);
            }
        }
        this->m_iNumPrevDeaths = 0;
        this->updateTotal();
        return BasePop::preWrite(fTime);
    }
    qhgb_pop *gpu() { return m_gpu; }
protected:
    qhgb_pop *m_gpu;
};
#endif
