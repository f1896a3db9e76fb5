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
#include "tut_EnvironAltPop.h"

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

struct RefSim {
    int nCells = 0;
    int nThreads = 1;
    bool quiet = true;
    SCellGrid *cg = nullptr;
    Geography *geo = nullptr;
    PopLooper *looper = nullptr;
    IDGen **idg = nullptr;
    tut_EnvironAltPop *pop = nullptr;
    uint32_t state[16];
    uint seeds[8];
    double tInit = 0, tActions = 0, tFinal = 0;
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

    s->pop = new tut_EnvironAltPop(s->cg, s->looper, layerSize > 0 ? layerSize : 65536, s->idg, s->state, s->seeds);
    ParamProvider2 *pp = ParamProvider2::createInstance(xml_path);
    int rc = -1;
    if (pp != NULL) {
        rc = pp->selectClass(class_name);
        if (rc == 0) rc = s->pop->readSpeciesData(pp);
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
        int start = s->pop->reserveAgentSpace(chunk);
#pragma omp parallel for
        for (int i = 0; i < chunk; i++) {
            tut_EnvironAltAgent &a = s->pop->m_aAgents[start + i];
            long j = done + i;
            a.m_iLifeState = life ? life[j] : LIFE_STATE_ALIVE;
            a.m_iCellIndex = cell[j];
            a.m_ulID = id[j];
            a.m_ulCellID = s->cg->m_aCells[cell[j]].m_iGlobalID;
            a.m_fBirthTime = birth[j];
            a.m_iGender = gender[j];
            a.m_fAge = age[j];
            a.m_fLastBirth = lastBirth[j];
            a.m_iMateIndex = -3;
        }
        for (int i = 0; i < chunk; i++) {
            if (id[done + i] > s->pop->m_iMaxID) s->pop->m_iMaxID = id[done + i];
        }
        done += chunk;
    }
    return 0;
}

int qref_start(void *h) {
    RefSim *s = (RefSim *)h;
    Quiet q(s->quiet);
    int rc = s->looper->addPop(s->pop);
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
        as += (int64_t)s->pop->getNumAgentsEffective();
        s->looper->doStep(t0 + i);
    }
    double w1 = omp_get_wtime();
    if (agentSteps) *agentSteps = as;
    return w1 - w0;
}

long qref_num_agents(void *h) { return (long)((RefSim *)h)->pop->getNumAgentsEffective(); }

// live agents in slot order; returns number written (or needed if cap too small)
long qref_get_agents(void *h, long cap, int *cell, int64_t *id, float *birth, uint8_t *gender, float *age,
                     float *lastBirth, uint32_t *life, int *mate, int *slot) {
    RefSim *s = (RefSim *)h;
    int first = s->pop->getFirstAgentIndex();
    if (first < 0) return 0;
    int last = s->pop->getLastAgentIndex();
    long k = 0;
    for (int i = first; i <= last; i++) {
        tut_EnvironAltAgent &a = s->pop->m_aAgents[i];
        if (a.m_iLifeState == LIFE_STATE_DEAD) continue;
        if (k < cap) {
            if (cell) cell[k] = a.m_iCellIndex;
            if (id) id[k] = a.m_ulID;
            if (birth) birth[k] = a.m_fBirthTime;
            if (gender) gender[k] = a.m_iGender;
            if (age) age[k] = a.m_fAge;
            if (lastBirth) lastBirth[k] = a.m_fLastBirth;
            if (life) life[k] = a.m_iLifeState;
            if (mate) mate[k] = a.m_iMateIndex;
            if (slot) slot[k] = i;
        }
        k++;
    }
    return k;
}

int qref_get_counts(void *h, uint64_t *out) {
    RefSim *s = (RefSim *)h;
    for (int c = 0; c < s->nCells; c++) out[c] = s->pop->getNumAgents(c);
    return 0;
}

int qref_get_weights(void *h, double *out) {  // nCells*7, actions/SingleEvaluator.cpp:174-243
    RefSim *s = (RefSim *)h;
    memcpy(out, s->pop->m_adEnvWeights, sizeof(double) * (size_t)s->nCells * 7);
    return 0;
}

int qref_get_bd(void *h, double *b, double *d) {  // actions/LinearBirth.cpp:97-112, LinearDeath.cpp:101-119
    RefSim *s = (RefSim *)h;
    if (s->pop->m_pVerhulst->m_pLB == NULL || s->pop->m_pVerhulst->m_pLD == NULL) return -1;
    memcpy(b, s->pop->m_pVerhulst->m_pLB->m_adB, sizeof(double) * s->nCells);
    memcpy(d, s->pop->m_pVerhulst->m_pLD->m_adD, sizeof(double) * s->nCells);
    return 0;
}

// ATanDeath probability exactly as actions/ATanDeath.cpp:49-59,75 computes it (fAge is the agent's float age)
int qref_atan_prob(void *h, int n, const float *age, double *p) {
    RefSim *s = (RefSim *)h;
    ATanDeath<tut_EnvironAltAgent> *ad = s->pop->m_pAD;
    for (int i = 0; i < n; i++) {
        p[i] = 0.5 + ad->m_dScale * atan(ad->m_dSlope * (age[i] - ad->m_dMaxAge)) / Q_PI;
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
    int rc = s->pop->updateEvent(EVENT_ID_GEO, NULL, t);
    s->pop->flushEvents(t);
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
    s->cg->delGeography();
    delete s->cg;
    delete s;
}

}  // extern "C"
