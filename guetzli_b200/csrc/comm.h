// Exchange step of the row-strip mode (one image tiled over the GPUs of a node,
// BASELINE configs[3]).  Every rank holds the full coefficient state
// and renders its own strip plus a 56-row halo itself (the metric's receptive
// field), so the only data that must cross NVLink are per-block
// results: the zeroing-order lists once, and one float per block per iteration
// (the per-block maxima of the distance map, from which every rank derives the
// same global distance).  Both are in-place all-gathers of uneven segments.
#pragma once
#include <stddef.h>

#include <vector>

#include "backend.h"

namespace gb200 {

class Comm {
 public:
  virtual ~Comm() {}
  virtual int rank() const = 0;
  virtual int world() const = 0;
  // Segment r (offset[r], count[r], in elements of elem_bytes) of dev_buf is valid
  // on rank r on entry; on return every segment is valid on every rank.  Ordered on
  // stream s.
  virtual void allgather_inplace(void* dev_buf, size_t elem_bytes, const std::vector<size_t>& offset,
                                 const std::vector<size_t>& count, Stream s) = 0;
};

// Block rows [lo, hi) owned by `rank` out of `bh` block rows.
inline void strip_of(int bh, int rank, int world, int* lo, int* hi) {
  *lo = static_cast<int>(static_cast<long long>(bh) * rank / world);
  *hi = static_cast<int>(static_cast<long long>(bh) * (rank + 1) / world);
}

}  // namespace gb200
