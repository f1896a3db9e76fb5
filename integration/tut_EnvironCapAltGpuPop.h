// populations/tut_EnvironCapAltGpuPop.h -- tut_EnvironCapAltPop (NPPCapacity, MultiEvaluator[NPP+Alt], VerhulstVarK) stepping on the GPU
#ifndef __TUT_ENVIRONCAPALTPOP_H__  // the header's own guard tests a misspelt macro (populations/tut_EnvironCapAltPop.h:1-2): it is not idempotent
#include "tut_EnvironCapAltPop.h"
#endif
#include "qhg_gpu_pop.h"
class tut_EnvironCapAltGpuPop : public QhgGpuPop<tut_EnvironCapAltPop, tut_EnvironCapAltAgent> {
public:
    tut_EnvironCapAltGpuPop(SCellGrid *pCG, PopFinder *pPF, int iLayerSize, IDGen **apIDG, uint32_t *aulState, uint *aiSeeds)
        : QhgGpuPop<tut_EnvironCapAltPop, tut_EnvironCapAltAgent>("tut_EnvironCapAltPop", pCG, pPF, iLayerSize, apIDG, aulState, aiSeeds) {}
};
