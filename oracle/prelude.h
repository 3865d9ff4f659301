// Force-included (-include) ahead of the reference sources in the ORACLE build only.
// The reference is MSVC code: it uses the MSVC-only sized-integer literal suffixes
// `0xffui8` / `0xffffui16` (NCC/NCC.h:7, SAD/Sad.h:10-11).  Under g++ these parse as
// user-defined literals, so two literal operators make the unmodified sources compile.
// It also leans on MSVC's <opencv2/opencv.hpp> transitively including <cstring>/<cstdlib>
// (cross_aggregator.cpp:108 memcpy, cross_aggregator.h:79 abs).
#pragma once
#ifdef __cplusplus
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <cassert>
constexpr unsigned char operator""ui8(unsigned long long v) { return static_cast<unsigned char>(v); }
constexpr unsigned short operator""ui16(unsigned long long v) { return static_cast<unsigned short>(v); }
#endif
