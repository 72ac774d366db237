// Host-side narrowing of count diagonals for the upload path (see hp_hostpack.cpp).
#pragma once
#include <cstddef>
#include <cstdint>

namespace hp {

// Writes src[0..len) to dst in the narrowest of u8 / u16 / i32 that holds every value exactly (counts are
// non-negative; a negative value forces i32).  dst must have room for len * 4 bytes.  Returns the element size
// chosen: 1, 2 or 4.
int narrow_diagonal(const int32_t* src, size_t len, void* dst);

// memcpy into a staging buffer that this core will not read again (the copy engine will): non-temporal stores, so the
// destination lines are not read before they are overwritten.  Any alignment, any size.
void stream_copy(void* dst, const void* src, size_t bytes);

}  // namespace hp
