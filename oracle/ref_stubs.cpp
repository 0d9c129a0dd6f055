// Link-time stubs for the libpng / CPython entry points the reference's asset I/O code
// references.  TEST INFRASTRUCTURE.  The oracle harness never loads or saves files, so
// each stub aborts loudly if it is ever reached.
#include <cstdio>
#include <cstdlib>
#include "png.h"
#include "Python.h"

[[noreturn]] static void unreachable_stub(const char *name) {
    std::fprintf(stderr, "oracle/_ref: stub '%s' called - asset I/O is not part of the oracle\n", name);
    std::abort();
}
#define STUB(name) unreachable_stub(#name)

extern "C" {
png_structp png_create_write_struct(const char *, void *, void *, void *) { STUB(png_create_write_struct); }
png_structp png_create_read_struct(const char *, void *, void *, void *) { STUB(png_create_read_struct); }
png_infop png_create_info_struct(png_structp) { STUB(png_create_info_struct); }
void png_destroy_write_struct(png_structp *, png_infop *) { STUB(png_destroy_write_struct); }
void png_destroy_read_struct(png_structp *, png_infop *, png_infop *) { STUB(png_destroy_read_struct); }
jmp_buf *png_shim_jmpbuf(png_structp) { STUB(png_jmpbuf); }
void png_init_io(png_structp, FILE *) { STUB(png_init_io); }
void png_set_IHDR(png_structp, png_infop, png_uint_32, png_uint_32, int, int, int, int, int) { STUB(png_set_IHDR); }
void png_write_info(png_structp, png_infop) { STUB(png_write_info); }
void png_write_row(png_structp, const png_byte *) { STUB(png_write_row); }
void png_write_end(png_structp, png_infop) { STUB(png_write_end); }
void png_read_info(png_structp, png_infop) { STUB(png_read_info); }
png_uint_32 png_get_image_width(png_structp, png_infop) { STUB(png_get_image_width); }
png_uint_32 png_get_image_height(png_structp, png_infop) { STUB(png_get_image_height); }
png_byte png_get_color_type(png_structp, png_infop) { STUB(png_get_color_type); }
png_byte png_get_bit_depth(png_structp, png_infop) { STUB(png_get_bit_depth); }
void png_set_strip_16(png_structp) { STUB(png_set_strip_16); }
void png_set_palette_to_rgb(png_structp) { STUB(png_set_palette_to_rgb); }
void png_set_expand_gray_1_2_4_to_8(png_structp) { STUB(png_set_expand_gray_1_2_4_to_8); }
png_uint_32 png_get_valid(png_structp, png_infop, png_uint_32) { STUB(png_get_valid); }
void png_set_tRNS_to_alpha(png_structp) { STUB(png_set_tRNS_to_alpha); }
void png_set_filler(png_structp, png_uint_32, int) { STUB(png_set_filler); }
void png_set_gray_to_rgb(png_structp) { STUB(png_set_gray_to_rgb); }
void png_set_strip_alpha(png_structp) { STUB(png_set_strip_alpha); }
void png_read_update_info(png_structp, png_infop) { STUB(png_read_update_info); }
void png_read_row(png_structp, png_bytep, png_bytep) { STUB(png_read_row); }

PyObject *Py_shim_None = nullptr;
void Py_Initialize(void) { STUB(Py_Initialize); }
int Py_IsInitialized(void) { return 0; }
void Py_Finalize(void) { STUB(Py_Finalize); }
PyObject *PySys_GetObject(const char *) { STUB(PySys_GetObject); }
PyObject *PyUnicode_DecodeFSDefault(const char *) { STUB(PyUnicode_DecodeFSDefault); }
PyObject *PyUnicode_FromString(const char *) { STUB(PyUnicode_FromString); }
int PyList_Append(PyObject *, PyObject *) { STUB(PyList_Append); }
PyObject *PyImport_Import(PyObject *) { STUB(PyImport_Import); }
PyObject *PyObject_GetAttrString(PyObject *, const char *) { STUB(PyObject_GetAttrString); }
int PyCallable_Check(PyObject *) { STUB(PyCallable_Check); }
PyObject *PyTuple_Pack(Py_ssize_t, ...) { STUB(PyTuple_Pack); }
PyObject *PyObject_CallObject(PyObject *, PyObject *) { STUB(PyObject_CallObject); }
PyObject *PyTuple_GetItem(PyObject *, Py_ssize_t) { STUB(PyTuple_GetItem); }
long PyLong_AsLong(PyObject *) { STUB(PyLong_AsLong); }
int PyObject_GetBuffer(PyObject *, Py_buffer *, int) { STUB(PyObject_GetBuffer); }
void PyBuffer_Release(Py_buffer *) { STUB(PyBuffer_Release); }
void PyErr_Print(void) { STUB(PyErr_Print); }
void Py_shim_decref(PyObject *) { STUB(Py_DECREF); }
}
