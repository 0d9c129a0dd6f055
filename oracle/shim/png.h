// Declaration-only stand-in for <png.h>.  TEST INFRASTRUCTURE.
// The reference's PNG load/save code compiles against this and links against the
// aborting stubs in oracle/ref_stubs.cpp; the oracle never reads or writes PNGs.
#pragma once
#include <csetjmp>
#include <cstdio>
#include <cstring>   // the reference's image.cpp uses std::memcpy without including it
typedef unsigned char png_byte;
typedef png_byte *png_bytep;
typedef struct png_struct_def png_struct;
typedef png_struct *png_structp;
typedef struct png_info_def png_info;
typedef png_info *png_infop;
typedef unsigned int png_uint_32;
#define PNG_LIBPNG_VER_STRING "shim"
#define PNG_COLOR_TYPE_GRAY 0
#define PNG_COLOR_TYPE_PALETTE 3
#define PNG_COLOR_TYPE_RGB 2
#define PNG_COLOR_TYPE_RGBA 6
#define PNG_COLOR_TYPE_GRAY_ALPHA 4
#define PNG_INTERLACE_NONE 0
#define PNG_COMPRESSION_TYPE_DEFAULT 0
#define PNG_FILTER_TYPE_DEFAULT 0
#define PNG_FILLER_AFTER 1
#define PNG_INFO_tRNS 0x0010U
extern "C" {
png_structp png_create_write_struct(const char *, void *, void *, void *);
png_structp png_create_read_struct(const char *, void *, void *, void *);
png_infop png_create_info_struct(png_structp);
void png_destroy_write_struct(png_structp *, png_infop *);
void png_destroy_read_struct(png_structp *, png_infop *, png_infop *);
jmp_buf *png_shim_jmpbuf(png_structp);
#define png_jmpbuf(p) (*png_shim_jmpbuf(p))
void png_init_io(png_structp, FILE *);
void png_set_IHDR(png_structp, png_infop, png_uint_32, png_uint_32, int, int, int, int, int);
void png_write_info(png_structp, png_infop);
void png_write_row(png_structp, const png_byte *);
void png_write_end(png_structp, png_infop);
void png_read_info(png_structp, png_infop);
png_uint_32 png_get_image_width(png_structp, png_infop);
png_uint_32 png_get_image_height(png_structp, png_infop);
png_byte png_get_color_type(png_structp, png_infop);
png_byte png_get_bit_depth(png_structp, png_infop);
void png_set_strip_16(png_structp);
void png_set_palette_to_rgb(png_structp);
void png_set_expand_gray_1_2_4_to_8(png_structp);
png_uint_32 png_get_valid(png_structp, png_infop, png_uint_32);
void png_set_tRNS_to_alpha(png_structp);
void png_set_filler(png_structp, png_uint_32, int);
void png_set_gray_to_rgb(png_structp);
void png_set_strip_alpha(png_structp);
void png_read_update_info(png_structp, png_infop);
void png_read_row(png_structp, png_bytep, png_bytep);
}
