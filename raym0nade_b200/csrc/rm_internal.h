// rm_internal.h — declarations shared by the translation units of libraym0nade_b200.
#pragma once
#include <cstdarg>
#include <cstdio>

// Records a printf-style message for rm_last_error() and returns `code`.
int rm_fail(int code, const char *fmt, ...);
