// dynpops/tut_EnvironAltGpuPopWrapper.cpp -- the plugin the reference's DynPopFactory loads (populations/DynPopFactory.cpp:80-163):
// a shared object whose name ends in "Wrapper.so", found in --so-dirs, exporting getInfo and createPop
// (dynpops/WrapperTemplate.cpp.tmp:14-31).  Select it with <class name="tut_EnvironAltGpuPop"> in the parameter file.
#include "PopBase.h"
#include "SCellGrid.h"
#include "IDGen.h"
#include "ArrayShare.h"
#include "PopFinder.h"
#include "tut_EnvironAltPop.cpp"   // the population's templates are compiled into the plugin, as in the reference's dynpops build
#include "tut_EnvironAltGpuPop.h"

extern "C" {
const std::string getInfo() { return "tut_EnvironAltGpuPop"; }

// DynPopFactory passes six arguments (populations/DynPopFactory.cpp:18,150); the seventh of the template is never read
PopBase *createPop(ArrayShare *pAS, SCellGrid *pCG, PopFinder *pPopFinder, int iLayerSize, IDGen **apIDG, uint32_t *aulState) {
    ArrayShare::setInstance(pAS);  // the plugin has its own statics: the application's ArrayShare is injected
    static uint aiNoSeeds[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    return new tut_EnvironAltGpuPop(pCG, pPopFinder, iLayerSize, apIDG, aulState, aiNoSeeds);
}
}
