/* TEST INFRASTRUCTURE ONLY -- never linked into the product library.
 *
 * Force-included (-include) ahead of the reference's
 *   /root/reference/DXRVoxelizer/XUSG/Optional/XUSGObjLoader.cpp
 * which is written against MSVC's precompiled stdafx.h and the *_s CRT calls.  This header
 * supplies the std headers stdafx.h would have provided and maps the three *_s calls the file
 * uses (fopen_s :21, fscanf_s :84.., sscanf_s :93..) to their ISO C counterparts.  The extra
 * buffer-size argument MSVC's fscanf_s("%s", buf, size) takes is an ignored surplus vararg for
 * ISO fscanf.  Nothing of the reference is copied: the .cpp is compiled where it lies.
 */
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

static inline int fopen_s(FILE** pp, const char* name, const char* mode)
{
    *pp = std::fopen(name, mode);
    return *pp ? 0 : 1;
}
#define fscanf_s fscanf
#define sscanf_s sscanf
