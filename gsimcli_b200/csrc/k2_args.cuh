// k2_args.cuh -- launch arguments shared by the two full-grid K2 kernels
// (k2_grid_sort.cu: register bitonic sort; k2_grid_csort.cu: counting sort)
#pragma once
#include "gsb_common.cuh"

struct K2Args {
    int R;
    int nchunks;
    int nplanes;
    int lockstep;      // experiment knob: block barriers inside the sort (GSB_K2_LOCKSTEP)
    int skew;          // warps 4..7 run the in-lane sort before the tile barrier (GSB_K2_SKEW)
    float nodata;
    long long ntiles;
    long long n0;      // first column of tile 0 (multiple of 4: TMA needs 16-byte aligned boxes)
    long long nn;      // columns from n0 to the end of the requested range
    long long skip;    // leading columns of tile 0 that precede the requested range
    float* out;        // plane 0, column n0 (may point `skip` elements before the buffer)
    long long out_pitch;
    const int* tile_list;   // counting-sort kernel only: [0] count, [1 + k] tile (NULL = all tiles)
    unsigned char kind[GSB_MAX_PLANES];
    double q[GSB_MAX_PLANES];
};

// per-warp scratch region of the counting-sort kernel, in 32-bit words:
// left halo, NPAD slots + one padding word per IPT slots, 2 x 32 atom bins and
// the missing-value bin (aliased by the right halo); multiples of 4 so the
// region can be cleared with 16-byte stores
#define K2_CS_LPAD(HALO) (((2 * (HALO)) + 3) & ~3)
#define K2_CS_REGION_WORDS(IPT, HALO) \
    ((K2_CS_LPAD(HALO) + (IPT) * 32 + 32 + 72 + 2 * (HALO) + 3) & ~3)

int gsb_launch_k2_csort(gsb_stack* s, const K2Args& args, int ipt, int grid, bool mode,
                        int halo, cudaStream_t st);
// register-tile kernel (k2_grid_rsel.cu); tiles it cannot finish exactly are
// appended to d_flagged ([0] count, [1 + k] tile) for the counting-sort kernel
int gsb_launch_k2_rsel(gsb_stack* s, const K2Args& args, int grid, bool mode, int nw,
                       int* d_flagged, cudaStream_t st);
// staged-select kernel (k2_grid_ssel.cu, experimental, k2_algo = 4): the reference's call
// shape only (gsb_k2_ssel_supports); flagged tiles as for the register-tile kernel
bool gsb_k2_ssel_supports(const K2Args& args, bool mode);
int gsb_launch_k2_ssel(gsb_stack* s, const K2Args& args, int grid, bool mode, int nw,
                       int* d_flagged, cudaStream_t st);
