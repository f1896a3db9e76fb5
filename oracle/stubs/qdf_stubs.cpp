// Link-time stand-ins for the reference's io/QDFUtils + io/WELLDumpRestore entry
// points that the population translation units reference but the step loop never
// calls (QDF = HDF5 file I/O, unavailable here).  TEST INFRASTRUCTURE ONLY.
// Signatures come from the reference headers included below (io/QDFUtils.h:157-226,
// io/WELLDumpRestore.h:9-10); bodies are failure/no-op.
#include <string>
#include "hdf5.h"
#include "QDFUtils.h"
#include "WELL512.h"
#include "WELLDumpRestore.h"

hid_t qdf_openGroup(hid_t, const std::string, bool) { return -1; }
hid_t qdf_opencreateGroup(hid_t, const std::string, bool) { return -1; }
void  qdf_closeGroup(hid_t) {}
void  qdf_closeDataSet(hid_t) {}
void  qdf_closeDataSpace(hid_t) {}
void  qdf_closeDataType(hid_t) {}
void  qdf_closeAttribute(hid_t) {}
bool  qdf_link_exists(hid_t, const std::string) { return false; }
int   qdf_insertSAttribute(hid_t, const std::string, const std::string) { return -1; }
int   qdf_extractSAttribute(hid_t, const std::string, std::string &) { return -1; }
int   qdf_insertAttribute(hid_t, const std::string, const uint, void *, const hid_t) { return -1; }
int   qdf_extractAttribute(hid_t, const std::string, const uint, void *, const hid_t) { return -1; }
PolyLine *qdf_createPolyLine(hid_t, const std::string) { return NULL; }
int   qdf_writePolyLine(hid_t, PolyLine *, const std::string) { return -1; }
int   qdf_compareDataTypes(hid_t, hid_t) { return -1; }
int   dumpWELL(WELL512 **, int, const std::string, hid_t) { return -1; }
int   restoreWELL(WELL512 **, int, const std::string, hid_t) { return -1; }
int   qdf_readArray(hid_t, const std::string, const uint, void *, const hid_t) { return -1; }
int   qdf_writeArray(hid_t, const std::string, const uint, void *, const hid_t) { return -1; }
int   qdf_replaceArray(hid_t, const std::string, const uint, void *, const hid_t) { return -1; }
