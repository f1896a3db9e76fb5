// populations/OoANavGenGpuPop.h -- OoANavGenPop (Genetics, Navigate, NPPCapacity ...: BASELINE configs #3 / #5) stepping on the GPU
#include "OoANavGenPop.h"
#include "qhg_gpu_pop.h"
class OoANavGenGpuPop : public QhgGpuPop<OoANavGenPop, OoANavGenAgent> {
public:
    OoANavGenGpuPop(SCellGrid *pCG, PopFinder *pPF, int iLayerSize, IDGen **apIDG, uint32_t *aulState, uint *aiSeeds)
        : QhgGpuPop<OoANavGenPop, OoANavGenAgent>("OoANavGenPop", pCG, pPF, iLayerSize, apIDG, aulState, aiSeeds) {}
};
