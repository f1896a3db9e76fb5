// populations/tut_EnvironAltGpuPop.h
#include "tut_EnvironAltPop.h"
#include "LBController.h"
extern "C" {
#include "qhg_b200.h"
}

class tut_EnvironAltGpuPop : public tut_EnvironAltPop {
public:
    tut_EnvironAltGpuPop(SCellGrid *pCG, PopFinder *pPF, int iLayerSize, IDGen **apIDG, uint32_t *aulState, uint *aiSeeds)
        : tut_EnvironAltPop(pCG, pPF, iLayerSize, apIDG, aulState, aiSeeds), m_gpu(NULL) {
        int iDev = getenv("QHG_DEVICE") ? atoi(getenv("QHG_DEVICE")) : 0;
        if (qhgb_create("tut_EnvironAltPop", iDev, pCG->m_iNumCells, pCG->m_iConnectivity, 0, &m_gpu) != 0) {
            xha_printf("[tut_EnvironAltGpuPop] %s\n", qhgb_last_error());   // no CPU fallback: the pop is unusable
        } else {
            qhgb_set_seed(m_gpu, aulState);
        }
    }
    virtual ~tut_EnvironAltGpuPop() { qhgb_destroy(m_gpu); }

    // XML parameters: the base class fills the Action objects; the same strings go to the device
    virtual int readSpeciesData(ParamProvider2 *pPP) {
        int iResult = tut_EnvironAltPop::readSpeciesData(pPP);
        const classinfo *pCI = pPP->getClassInfo();
        for (stringmap::const_iterator it = pCI->prios.begin(); iResult == 0 && it != pCI->prios.end(); ++it)
            iResult += qhgb_set_prio(m_gpu, it->first.c_str(), atoi(it->second.c_str()));
        for (modulemap::const_iterator im = pCI->mods.begin(); iResult == 0 && im != pCI->mods.end(); ++im)
            iResult += pushParams(im->second);
        return iResult;
    }

    // parameter name -> value string of a module and of its sub-modules (the evaluators inside a MultiEvaluator)
    int pushParams(const ModuleComplex *pM) {
        int iResult = 0;
        const stringmap &mA = pM->getAttributes();
        for (stringmap::const_iterator ip = mA.begin(); ip != mA.end(); ++ip)
            iResult += qhgb_set_attribute_str(m_gpu, ip->first.c_str(), ip->second.c_str());
        const modulemap &mS = pM->getSubModules();
        for (modulemap::const_iterator is = mS.begin(); is != mS.end(); ++is) iResult += pushParams(is->second);
        return iResult;
    }

    virtual int preLoop() {
        int iResult = tut_EnvironAltPop::preLoop();
        // grid + environment
        std::vector<int32_t> vN(m_iNumCells * 6), vID(m_iNumCells);
        std::vector<double> vIce(m_iNumCells);
        for (int c = 0; c < m_iNumCells; c++) {
            vID[c] = m_pCG->m_aCells[c].m_iGlobalID;
            for (int k = 0; k < 6; k++) vN[6*c + k] = m_pCG->m_aCells[c].m_aNeighbors[k];
            vIce[c] = m_pGeography->m_abIce[c];
        }
        iResult += qhgb_set_cells(m_gpu, vN.data(), vID.data());
        iResult += qhgb_set_env_array(m_gpu, "Altitude", m_pGeography->m_adAltitude, m_iNumCells);
        iResult += qhgb_set_env_array(m_gpu, "Ice", vIce.data(), m_iNumCells);
        // agents loaded by addAgent / readAgentDataQDF -> structure of arrays
        std::vector<int32_t> c; std::vector<int64_t> id; std::vector<float> b, a, l; std::vector<uint8_t> g; std::vector<uint32_t> s;
        int i0 = getFirstAgentIndex();
        if (i0 != LBController::NIL) for (int i = i0; i <= getLastAgentIndex(); i++) {
            tut_EnvironAltAgent &ag = m_aAgents[i];
            if (ag.m_iLifeState == LIFE_STATE_DEAD) continue;
            c.push_back(ag.m_iCellIndex); id.push_back(ag.m_ulID); b.push_back(ag.m_fBirthTime); g.push_back(ag.m_iGender);
            a.push_back(ag.m_fAge); l.push_back(ag.m_fLastBirth); s.push_back(ag.m_iLifeState);
        }
        iResult += qhgb_add_agents(m_gpu, c.size(), c.data(), id.data(), b.data(), g.data(), a.data(), l.data(), s.data());
        iResult += qhgb_pre_loop(m_gpu);
        return iResult;
    }

    virtual int initializeStep(float fTime) { m_fCurTime = fTime; return qhgb_initialize_step(m_gpu, fTime); }
    virtual int doActions(uint iPrio, float fTime) { return qhgb_do_actions(m_gpu, iPrio, fTime); }
    virtual int finalizeStep() {
        int iResult = qhgb_finalize_step(m_gpu);
        qhgb_step_stats st;
        qhgb_get_step_stats(m_gpu, &st);
        m_iNumBirths += st.births; m_iNumDeaths += st.deaths; m_iNumMoves += st.moves;
        return iResult;
    }
    virtual ulong getNumAgentsEffective() { return qhgb_get_num_agents_effective(m_gpu); }
    virtual ulong getNumAgentsTotal()     { return qhgb_get_num_agents_effective(m_gpu); }
    virtual void  updateNumAgentsPerCell() { qhgb_get_num_agents_array(m_gpu, (uint64_t *)m_aiNumAgentsPerCell); }
    virtual ulong getNumAgents(int iCell)  { updateNumAgentsPerCell(); return m_aiNumAgentsPerCell[iCell]; }

    virtual int updateEvent(int iEventID, char *pData, float fT) {     // arrays were re-read by Simulator::handleEnvironmentEvent
        std::vector<double> vIce(m_iNumCells);
        for (int c = 0; c < m_iNumCells; c++) vIce[c] = m_pGeography->m_abIce[c];
        int iResult = qhgb_set_env_array(m_gpu, "Altitude", m_pGeography->m_adAltitude, m_iNumCells);
        iResult += qhgb_set_env_array(m_gpu, "Ice", vIce.data(), m_iNumCells);
        iResult += qhgb_update_event(m_gpu, iEventID, fT);
        notifyObservers(iEventID, pData);
        return iResult;
    }
    virtual void flushEvents(float fT) { qhgb_flush_events(m_gpu, fT); notifyObservers(EVENT_ID_FLUSH, NULL); }

    // the host calls PopLooper::preWrite right before every QDF write / dump (app/Simulator.cpp:601): bring the device
    // state back into m_aAgents so that the unchanged writeAgentDataQDFSafe works
    virtual int preWrite(float fTime) {
        int64_t n = qhgb_get_num_agents_effective(m_gpu);
        std::vector<int32_t> c(n), cid(n); std::vector<int64_t> id(n); std::vector<float> b(n), a(n), l(n);
        std::vector<uint8_t> g(n); std::vector<uint32_t> s(n);
        if (qhgb_get_agents(m_gpu, n, c.data(), cid.data(), id.data(), b.data(), g.data(), a.data(), l.data(), s.data(), NULL) != n) return -1;
        m_pAgentController->clear();
        int iStart = n > 0 ? reserveAgentSpace((int)n) : 0;
        for (int64_t i = 0; i < n; i++) {
            tut_EnvironAltAgent &ag = m_aAgents[iStart + i];
            ag.m_iLifeState = s[i]; ag.m_iCellIndex = c[i]; ag.m_ulCellID = cid[i]; ag.m_ulID = id[i];
            ag.m_fBirthTime = b[i]; ag.m_iGender = g[i]; ag.m_fAge = a[i]; ag.m_fLastBirth = l[i]; ag.m_iMateIndex = -3;
        }
        m_iNumPrevDeaths = 0;
        updateTotal();
        return tut_EnvironAltPop::preWrite(fTime);
    }
protected:
    qhgb_pop *m_gpu;
};
