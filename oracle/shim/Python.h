// Declaration-only stand-in for <Python.h>.  TEST INFRASTRUCTURE.
// The reference decodes DDS/HDR assets through embedded CPython (material.cpp:150-218,
// component.cpp:69-114).  The oracle injects synthetic textures directly, so these
// entry points are linked against aborting stubs (oracle/ref_stubs.cpp) and never run.
#pragma once
// The real Python.h pulls in the C headers below (pyport.h includes <math.h>).  That is
// load-bearing for the reference: libstdc++'s <math.h> wrapper injects the float overloads
// (`using std::sqrt;` ...) into the global namespace, so the reference's unqualified
// sqrt/cos/pow/log2/isnan calls on floats resolve to the float versions in every
// translation unit that includes material.h.  (geometry.cpp does not, and there the same
// calls resolve to the double versions.)  The shim must reproduce that.
#include <stdio.h>
#include <string.h>
#include <errno.h>
#include <stdlib.h>
#include <stddef.h>
#include <assert.h>
#include <math.h>
typedef struct _shim_pyobject PyObject;
typedef long Py_ssize_t;
typedef struct { void *buf; PyObject *obj; Py_ssize_t len; } Py_buffer;
#define PyBUF_SIMPLE 0
extern "C" {
extern PyObject *Py_shim_None;
#define Py_None Py_shim_None
void Py_Initialize(void);
int Py_IsInitialized(void);
void Py_Finalize(void);
PyObject *PySys_GetObject(const char *);
PyObject *PyUnicode_DecodeFSDefault(const char *);
PyObject *PyUnicode_FromString(const char *);
int PyList_Append(PyObject *, PyObject *);
PyObject *PyImport_Import(PyObject *);
PyObject *PyObject_GetAttrString(PyObject *, const char *);
int PyCallable_Check(PyObject *);
PyObject *PyTuple_Pack(Py_ssize_t, ...);
PyObject *PyObject_CallObject(PyObject *, PyObject *);
PyObject *PyTuple_GetItem(PyObject *, Py_ssize_t);
long PyLong_AsLong(PyObject *);
int PyObject_GetBuffer(PyObject *, Py_buffer *, int);
void PyBuffer_Release(Py_buffer *);
void PyErr_Print(void);
void Py_shim_decref(PyObject *);
}
#define Py_DECREF(o) Py_shim_decref((PyObject *)(o))
#define Py_XDECREF(o) Py_shim_decref((PyObject *)(o))
