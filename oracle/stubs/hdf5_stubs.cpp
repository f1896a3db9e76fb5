// Link-time stand-ins for the HDF5 C API (see hdf5.h in this directory).
// TEST INFRASTRUCTURE ONLY.  The per-step agent update never performs file I/O, so none of these is reached while stepping.
// The calls of the AGENT DATASET path -- what io/PopWriter.cpp:84-118 and io/PopReader.cpp:143-170 do around
// PopBase::writeAgentDataQDF / readAgentDataQDF (core/SPopulation.cpp:1356-1372,1405-1417,1465-1568,1689-1741): compound type,
// 1-D dataspaces with a hyperslab selection, one dataset -- are backed by memory, so that the reference's own writers and
// readers run here and the records they hand to HDF5 can be looked at (tests/test_qdf_agent_io.py).  Everything else (files,
// groups, attributes) reports failure (-1) or does nothing.
#include "hdf5.h"
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

namespace {
struct StubSpace { hsize_t dims = 0, off = 0, cnt = 0; bool sel = false; };
struct StubMember { std::string name; size_t offset; hid_t type; };
struct StubType { size_t size = 0; std::vector<StubMember> members; };
struct StubSet { size_t elem = 0; hsize_t n = 0; hid_t type = -1; std::vector<unsigned char> bytes; };
std::map<hid_t, StubSpace> g_spaces;
std::map<hid_t, StubType> g_types;
std::map<hid_t, StubSet> g_sets;
hid_t g_next = 1000;
size_t nativeSize(hid_t t) {
    switch ((int)t) {
    case QHGSTUB_T_CHAR: case QHGSTUB_T_UCHAR: case QHGSTUB_T_INT8: case QHGSTUB_T_UINT8: case QHGSTUB_T_HBOOL: case QHGSTUB_T_C_S1: return 1;
    case QHGSTUB_T_SHORT: case QHGSTUB_T_USHORT: case QHGSTUB_T_INT16: case QHGSTUB_T_UINT16: return 2;
    case QHGSTUB_T_INT: case QHGSTUB_T_UINT: case QHGSTUB_T_FLOAT: case QHGSTUB_T_INT32: case QHGSTUB_T_UINT32: return 4;
    case QHGSTUB_T_LONG: case QHGSTUB_T_ULONG: case QHGSTUB_T_LLONG: case QHGSTUB_T_ULLONG: case QHGSTUB_T_DOUBLE: case QHGSTUB_T_INT64:
    case QHGSTUB_T_UINT64: return 8;
    case QHGSTUB_T_LDOUBLE: return 16;
    default: return 0;
    }
}
}  // namespace

// what the tests look at (not HDF5 API): the bytes of a dataset, the members of a compound type
extern "C" long qhgstub_dataset_info(hid_t dset, size_t *elem, hid_t *type) {
    auto it = g_sets.find(dset);
    if (it == g_sets.end()) return -1;
    if (elem) *elem = it->second.elem;
    if (type) *type = it->second.type;
    return (long)it->second.n;
}
extern "C" int qhgstub_dataset_bytes(hid_t dset, void *out) {
    auto it = g_sets.find(dset);
    if (it == g_sets.end()) return -1;
    memcpy(out, it->second.bytes.data(), it->second.bytes.size());
    return 0;
}
extern "C" hid_t qhgstub_dataset_from_bytes(hid_t type, hsize_t n, const void *bytes) {  // a dataset "found in a file"
    auto it = g_types.find(type);
    if (it == g_types.end()) return -1;
    StubSet d; d.elem = it->second.size; d.n = n; d.type = type;
    d.bytes.assign((const unsigned char *)bytes, (const unsigned char *)bytes + d.elem * n);
    g_sets[g_next] = d;
    return g_next++;
}
extern "C" int qhgstub_type_member(hid_t type, int i, char *name, int cap, size_t *offset, hid_t *mtype) {
    auto it = g_types.find(type);
    if (it == g_types.end() || i < 0 || i >= (int)it->second.members.size()) return -1;
    const StubMember &m = it->second.members[i];
    snprintf(name, cap, "%s", m.name.c_str());
    *offset = m.offset; *mtype = m.type;
    return 0;
}

extern "C" {
hid_t   H5Screate_simple(int rank, const hsize_t *dims, const hsize_t *) { if (rank != 1) return -1; StubSpace s; s.dims = dims[0]; g_spaces[g_next] = s; return g_next++; }
hid_t   H5Screate(int) { return -1; }
herr_t  H5Sselect_hyperslab(hid_t sp, H5S_seloper_t, const hsize_t *start, const hsize_t *stride, const hsize_t *count, const hsize_t *block) {
    auto it = g_spaces.find(sp);
    if (it == g_spaces.end() || (stride && stride[0] != 1) || (block && block[0] != 1) || start[0] + count[0] > it->second.dims) return -1;
    it->second.off = start[0]; it->second.cnt = count[0]; it->second.sel = true;
    return 0;
}
int     H5Sget_simple_extent_dims(hid_t sp, hsize_t *dims, hsize_t *) { auto it = g_spaces.find(sp); if (it == g_spaces.end()) return -1; if (dims) dims[0] = it->second.dims; return 1; }
int     H5Sget_simple_extent_ndims(hid_t sp) { return g_spaces.count(sp) ? 1 : -1; }
herr_t  H5Sclose(hid_t sp) { g_spaces.erase(sp); return 0; }
hid_t   H5Tcreate(H5T_class_t cls, size_t size) { if (cls != H5T_COMPOUND) return 1; StubType t; t.size = size; g_types[g_next] = t; return g_next++; }
herr_t  H5Tinsert(hid_t t, const char *name, size_t off, hid_t m) { auto it = g_types.find(t); if (it != g_types.end()) it->second.members.push_back({name, off, m}); return 0; }
herr_t  H5Tset_size(hid_t, size_t) { return 0; }
hid_t   H5Tcopy(hid_t t) { return t; }
herr_t  H5Tclose(hid_t) { return 0; }
htri_t  H5Tequal(hid_t a, hid_t b) { return a == b; }
int     H5Tget_nmembers(hid_t t) { auto it = g_types.find(t); return it == g_types.end() ? -1 : (int)it->second.members.size(); }
size_t  H5Tget_member_offset(hid_t, unsigned) { return 0; }
hid_t   H5Tget_member_type(hid_t, unsigned) { return -1; }
char   *H5Tget_member_name(hid_t, unsigned) { return nullptr; }
int     H5Tget_member_index(hid_t, const char *) { return -1; }
size_t  H5Tget_size(hid_t t) { auto it = g_types.find(t); return it == g_types.end() ? nativeSize(t) : it->second.size; }
herr_t  H5free_memory(void *p) { free(p); return 0; }
herr_t  H5Dwrite(hid_t dset, hid_t memtype, hid_t memspace, hid_t filespace, hid_t, const void *buf) {
    auto d = g_sets.find(dset); auto f = g_spaces.find(filespace); auto m = g_spaces.find(memspace);
    if (d == g_sets.end() || f == g_spaces.end() || m == g_spaces.end() || H5Tget_size(memtype) != d->second.elem) return -1;
    const hsize_t off = f->second.sel ? f->second.off : 0, cnt = f->second.sel ? f->second.cnt : f->second.dims;
    if (cnt != m->second.dims || off + cnt > d->second.n) return -1;   // the memory space holds exactly the selected elements
    memcpy(d->second.bytes.data() + off * d->second.elem, buf, cnt * d->second.elem);
    return 0;
}
herr_t  H5Dread(hid_t dset, hid_t memtype, hid_t memspace, hid_t filespace, hid_t, void *buf) {
    auto d = g_sets.find(dset); auto f = g_spaces.find(filespace); auto m = g_spaces.find(memspace);
    if (d == g_sets.end() || f == g_spaces.end() || m == g_spaces.end() || H5Tget_size(memtype) != d->second.elem) return -1;
    const hsize_t off = f->second.sel ? f->second.off : 0, cnt = f->second.sel ? f->second.cnt : f->second.dims;
    if (cnt != m->second.dims || off + cnt > d->second.n) return -1;
    memcpy(buf, d->second.bytes.data() + off * d->second.elem, cnt * d->second.elem);
    return 0;
}
hid_t   H5Dopen2(hid_t, const char *, hid_t) { return -1; }
hid_t   H5Dget_space(hid_t dset) { auto d = g_sets.find(dset); if (d == g_sets.end()) return -1; StubSpace s; s.dims = d->second.n; g_spaces[g_next] = s; return g_next++; }
hid_t   H5Dget_type(hid_t) { return -1; }
hid_t   H5Dcreate2(hid_t, const char *, hid_t type, hid_t space, hid_t, hid_t, hid_t) {
    auto t = g_types.find(type); auto sp = g_spaces.find(space);
    if (t == g_types.end() || sp == g_spaces.end()) return -1;
    StubSet d; d.elem = t->second.size; d.n = sp->second.dims; d.type = type; d.bytes.assign(d.elem * d.n, 0);
    g_sets[g_next] = d;
    return g_next++;
}
herr_t  H5Dclose(hid_t) { return 0; }
herr_t  H5Awrite(hid_t, hid_t, const void *) { return -1; }
herr_t  H5Aread(hid_t, hid_t, void *) { return -1; }
hid_t   H5Aopen_name(hid_t, const char *) { return -1; }
hid_t   H5Aopen(hid_t, const char *, hid_t) { return -1; }
hid_t   H5Aget_type(hid_t) { return -1; }
hid_t   H5Aget_space(hid_t) { return -1; }
htri_t  H5Aexists(hid_t, const char *) { return 0; }
hid_t   H5Acreate(hid_t, const char *, hid_t, hid_t, hid_t, hid_t) { return -1; }
hid_t   H5Acreate2(hid_t, const char *, hid_t, hid_t, hid_t, hid_t) { return -1; }
herr_t  H5Aclose(hid_t) { return 0; }
hid_t   H5Gopen2(hid_t, const char *, hid_t) { return -1; }
hid_t   H5Gcreate2(hid_t, const char *, hid_t, hid_t, hid_t) { return -1; }
herr_t  H5Gclose(hid_t) { return 0; }
hid_t   H5Fopen(const char *, unsigned, hid_t) { return -1; }
hid_t   H5Fcreate(const char *, unsigned, hid_t, hid_t) { return -1; }
herr_t  H5Fclose(hid_t) { return 0; }
htri_t  H5Lexists(hid_t, const char *, hid_t) { return 0; }
}
