// Link-time stand-ins for the HDF5 C API (see hdf5.h in this directory).
// TEST INFRASTRUCTURE ONLY.  Every call reports failure (-1) or does nothing:
// the per-step agent update never performs file I/O, so none of these is
// reached while stepping; QDF read/write through oracle/_ref is unsupported.
#include "hdf5.h"
#include <cstdlib>

extern "C" {
hid_t   H5Screate_simple(int, const hsize_t *, const hsize_t *) { return -1; }
hid_t   H5Screate(int) { return -1; }
herr_t  H5Sselect_hyperslab(hid_t, H5S_seloper_t, const hsize_t *, const hsize_t *, const hsize_t *, const hsize_t *) { return -1; }
int     H5Sget_simple_extent_dims(hid_t, hsize_t *, hsize_t *) { return -1; }
int     H5Sget_simple_extent_ndims(hid_t) { return -1; }
herr_t  H5Sclose(hid_t) { return 0; }
hid_t   H5Tcreate(H5T_class_t, size_t) { return 1; }
herr_t  H5Tinsert(hid_t, const char *, size_t, hid_t) { return 0; }
herr_t  H5Tset_size(hid_t, size_t) { return 0; }
hid_t   H5Tcopy(hid_t t) { return t; }
herr_t  H5Tclose(hid_t) { return 0; }
htri_t  H5Tequal(hid_t a, hid_t b) { return a == b; }
int     H5Tget_nmembers(hid_t) { return -1; }
size_t  H5Tget_member_offset(hid_t, unsigned) { return 0; }
hid_t   H5Tget_member_type(hid_t, unsigned) { return -1; }
char   *H5Tget_member_name(hid_t, unsigned) { return nullptr; }
int     H5Tget_member_index(hid_t, const char *) { return -1; }
size_t  H5Tget_size(hid_t) { return 0; }
herr_t  H5free_memory(void *p) { free(p); return 0; }
herr_t  H5Dwrite(hid_t, hid_t, hid_t, hid_t, hid_t, const void *) { return -1; }
herr_t  H5Dread(hid_t, hid_t, hid_t, hid_t, hid_t, void *) { return -1; }
hid_t   H5Dopen2(hid_t, const char *, hid_t) { return -1; }
hid_t   H5Dget_space(hid_t) { return -1; }
hid_t   H5Dget_type(hid_t) { return -1; }
hid_t   H5Dcreate2(hid_t, const char *, hid_t, hid_t, hid_t, hid_t, hid_t) { return -1; }
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
