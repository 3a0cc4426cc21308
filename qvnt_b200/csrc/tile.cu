// tile.cu -- fused tile pass (placeholder until the kernel lands)
#include "engine.h"
namespace qv {
size_t tile_smem_bytes(uint32_t T) { return (size_t)16 << T; }
int launch_tile_pass(cudaStream_t, const Segs &, const TilePass &, const TileOp *, const amp *, int, int) { return -1; }
}
