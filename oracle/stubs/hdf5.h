/* Stand-in <hdf5.h> for building the reference's step loop WITHOUT libhdf5.
 *
 * TEST INFRASTRUCTURE ONLY (used by oracle/Makefile to compile the unmodified
 * reference sources under /root/reference into oracle/_ref/).  HDF5 is not
 * installed in this environment and the per-step agent update never touches a
 * file, so every call below is a declaration whose definition (hdf5_stubs.cpp)
 * fails or does nothing.  Written from the public HDF5 C API names; it contains
 * no reference code.
 */
#ifndef QHG_B200_ORACLE_HDF5_STUB_H
#define QHG_B200_ORACLE_HDF5_STUB_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef int64_t hid_t;
typedef int herr_t;
typedef int htri_t;
typedef unsigned long long hsize_t;
typedef long long hssize_t;
typedef int H5T_class_t;
typedef int H5S_seloper_t;
#define H5P_DEFAULT      ((hid_t)0)
#define H5S_ALL          ((hid_t)0)
#define H5S_SELECT_SET   0
#define H5S_SCALAR       0
#define H5S_SIMPLE       1
#define H5T_COMPOUND     6
#define H5T_STRING       3
#define H5F_ACC_RDONLY   0u
#define H5F_ACC_RDWR     1u
#define H5F_ACC_TRUNC    2u
#define H5T_VARIABLE     ((size_t)(-1))
#define H5_INDEX_NAME    0
#define H5_ITER_INC      0
#define H5_ITER_NATIVE   2
#define H5E_DEFAULT      ((hid_t)0)
#define HOFFSET(S, M)    (offsetof(S, M))
enum { QHGSTUB_T_CHAR = 101, QHGSTUB_T_UCHAR, QHGSTUB_T_SHORT, QHGSTUB_T_USHORT, QHGSTUB_T_INT,
       QHGSTUB_T_UINT, QHGSTUB_T_LONG, QHGSTUB_T_ULONG, QHGSTUB_T_LLONG, QHGSTUB_T_ULLONG,
       QHGSTUB_T_FLOAT, QHGSTUB_T_DOUBLE, QHGSTUB_T_INT32, QHGSTUB_T_UINT32, QHGSTUB_T_INT64,
       QHGSTUB_T_UINT64, QHGSTUB_T_C_S1, QHGSTUB_T_HBOOL, QHGSTUB_T_INT8, QHGSTUB_T_UINT8,
       QHGSTUB_T_LDOUBLE, QHGSTUB_T_INT16, QHGSTUB_T_UINT16 };
#define H5T_NATIVE_CHAR   ((hid_t)QHGSTUB_T_CHAR)
#define H5T_NATIVE_SCHAR  ((hid_t)QHGSTUB_T_CHAR)
#define H5T_NATIVE_UCHAR  ((hid_t)QHGSTUB_T_UCHAR)
#define H5T_NATIVE_SHORT  ((hid_t)QHGSTUB_T_SHORT)
#define H5T_NATIVE_USHORT ((hid_t)QHGSTUB_T_USHORT)
#define H5T_NATIVE_INT    ((hid_t)QHGSTUB_T_INT)
#define H5T_NATIVE_UINT   ((hid_t)QHGSTUB_T_UINT)
#define H5T_NATIVE_LONG   ((hid_t)QHGSTUB_T_LONG)
#define H5T_NATIVE_ULONG  ((hid_t)QHGSTUB_T_ULONG)
#define H5T_NATIVE_LLONG  ((hid_t)QHGSTUB_T_LLONG)
#define H5T_NATIVE_ULLONG ((hid_t)QHGSTUB_T_ULLONG)
#define H5T_NATIVE_FLOAT  ((hid_t)QHGSTUB_T_FLOAT)
#define H5T_NATIVE_DOUBLE ((hid_t)QHGSTUB_T_DOUBLE)
#define H5T_NATIVE_LDOUBLE ((hid_t)QHGSTUB_T_LDOUBLE)
#define H5T_NATIVE_INT16  ((hid_t)QHGSTUB_T_INT16)
#define H5T_NATIVE_UINT16 ((hid_t)QHGSTUB_T_UINT16)
#define H5T_NATIVE_INT32  ((hid_t)QHGSTUB_T_INT32)
#define H5T_NATIVE_UINT32 ((hid_t)QHGSTUB_T_UINT32)
#define H5T_NATIVE_INT64  ((hid_t)QHGSTUB_T_INT64)
#define H5T_NATIVE_UINT64 ((hid_t)QHGSTUB_T_UINT64)
#define H5T_NATIVE_INT8   ((hid_t)QHGSTUB_T_INT8)
#define H5T_NATIVE_UINT8  ((hid_t)QHGSTUB_T_UINT8)
#define H5T_NATIVE_HBOOL  ((hid_t)QHGSTUB_T_HBOOL)
#define H5T_C_S1          ((hid_t)QHGSTUB_T_C_S1)

/* dataspaces */
hid_t   H5Screate_simple(int rank, const hsize_t *dims, const hsize_t *maxdims);
hid_t   H5Screate(int type);
herr_t  H5Sselect_hyperslab(hid_t space, H5S_seloper_t op, const hsize_t *start, const hsize_t *stride,
                            const hsize_t *count, const hsize_t *block);
int     H5Sget_simple_extent_dims(hid_t space, hsize_t *dims, hsize_t *maxdims);
int     H5Sget_simple_extent_ndims(hid_t space);
herr_t  H5Sclose(hid_t space);
/* datatypes */
hid_t   H5Tcreate(H5T_class_t cls, size_t size);
herr_t  H5Tinsert(hid_t parent, const char *name, size_t offset, hid_t member);
herr_t  H5Tset_size(hid_t type, size_t size);
hid_t   H5Tcopy(hid_t type);
herr_t  H5Tclose(hid_t type);
htri_t  H5Tequal(hid_t a, hid_t b);
int     H5Tget_nmembers(hid_t type);
size_t  H5Tget_member_offset(hid_t type, unsigned membno);
hid_t   H5Tget_member_type(hid_t type, unsigned membno);
char   *H5Tget_member_name(hid_t type, unsigned membno);
int     H5Tget_member_index(hid_t type, const char *name);
size_t  H5Tget_size(hid_t type);
herr_t  H5free_memory(void *mem);
/* datasets */
herr_t  H5Dwrite(hid_t dset, hid_t memtype, hid_t memspace, hid_t filespace, hid_t plist, const void *buf);
herr_t  H5Dread(hid_t dset, hid_t memtype, hid_t memspace, hid_t filespace, hid_t plist, void *buf);
hid_t   H5Dopen2(hid_t loc, const char *name, hid_t dapl);
hid_t   H5Dget_space(hid_t dset);
hid_t   H5Dget_type(hid_t dset);
hid_t   H5Dcreate2(hid_t loc, const char *name, hid_t type, hid_t space, hid_t lcpl, hid_t dcpl, hid_t dapl);
herr_t  H5Dclose(hid_t dset);
/* attributes */
herr_t  H5Awrite(hid_t attr, hid_t type, const void *buf);
herr_t  H5Aread(hid_t attr, hid_t type, void *buf);
hid_t   H5Aopen_name(hid_t loc, const char *name);
hid_t   H5Aopen(hid_t loc, const char *name, hid_t aapl);
hid_t   H5Aget_type(hid_t attr);
hid_t   H5Aget_space(hid_t attr);
htri_t  H5Aexists(hid_t loc, const char *name);
hid_t   H5Acreate(hid_t loc, const char *name, hid_t type, hid_t space, hid_t acpl, hid_t aapl);
hid_t   H5Acreate2(hid_t loc, const char *name, hid_t type, hid_t space, hid_t acpl, hid_t aapl);
herr_t  H5Aclose(hid_t attr);
/* groups / files (never reached by the step loop) */
hid_t   H5Gopen2(hid_t loc, const char *name, hid_t gapl);
hid_t   H5Gcreate2(hid_t loc, const char *name, hid_t lcpl, hid_t gcpl, hid_t gapl);
herr_t  H5Gclose(hid_t group);
hid_t   H5Fopen(const char *name, unsigned flags, hid_t fapl);
hid_t   H5Fcreate(const char *name, unsigned flags, hid_t fcpl, hid_t fapl);
herr_t  H5Fclose(hid_t file);
htri_t  H5Lexists(hid_t loc, const char *name, hid_t lapl);
#ifdef __cplusplus
}
#endif
#endif
