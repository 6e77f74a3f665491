// patch.cuh -- interface of the multi-box / multi-level ("AMR-exact") kernels (patch.cu).
//
// A patch level is a list of boxes, each stored as an AMReX-shaped FAB: the valid box grown by NG = 3 ghost
// cells (m_f_nghost, Source/LBM.H:230) in every direction, x fastest, component slowest
// (AMReX_Array4.H:60-94) -- so a box can be library-owned memory or the device memory of an AMReX MultiFab FAB
// (mbl_level_bind, zero-copy).  The operators follow the reference's un-fused sequence exactly
// (Source/LBM.cpp:523-618): grown-box stream with the -1 sentinel, average-down between stream and collide,
// ghost cells of a fine level that stream twice between two coarse fills.
#pragma once
#include "kernels.cuh"

namespace mbl {

constexpr int PNG = 3;  // ghost cells of f, g, is_fluid on a patch level

struct PBox {
    int lo[3], hi[3];   // valid box, level index space
    int glo[3], n[3];   // corner and extents of the grown box
    long long sy, sz, sq;  // strides (doubles): row, plane, component
    double* f[2];       // lattice buffers; [cur] is the level's current state
    double* g[2];
    double* qc;         // 3 comps, same shape (QCorr of the post-stream state on the valid box grown by 1)
    double* macro;      // 19 + 7 comps, same shape, or nullptr
    int32_t* isfl;      // is_fluid comp 0, same shape
    __host__ __device__ long long cell(int i, int j, int k) const
    {
        return (long long)(i - glo[0]) + (long long)(j - glo[1]) * sy + (long long)(k - glo[2]) * sz;
    }
    __host__ __device__ bool in_grown(int i, int j, int k) const
    {
        return i >= glo[0] && i < glo[0] + n[0] && j >= glo[1] && j < glo[1] + n[1] && k >= glo[2] && k < glo[2] + n[2];
    }
    __host__ __device__ bool in_valid(int i, int j, int k) const
    {
        return i >= lo[0] && i <= hi[0] && j >= lo[1] && j <= hi[1] && k >= lo[2] && k <= hi[2];
    }
    __host__ __device__ long long ncell() const { return sq; }
};

struct PGeom {
    int dlo[3], dhi[3];  // level domain
    int periodic[3];
};

// one rectangular copy: n cells starting at d (in the destination box) from s (in the source box)
struct CopyTag {
    int dbox, sbox;
    int d[3], s[3], n[3];
    long long boff;  // remote tags: first cell of this tag in the peer's message (cells, not doubles)
};
// which array of a PBox a copy / fill touches
enum PArray { PA_F = 0, PA_G = 1, PA_QC = 2, PA_MACRO = 3 };

struct RegionTag {
    int box;
    int lo[3], n[3];
};

void patch_init_tables();

int launch_patch_copy(const PBox* dtab, int dcur, const PBox* stab, int scur, const CopyTag* tags, int ntags, int darr,
                      int sarr, int ncomp, long long max_cells, cudaStream_t st, bool skip_keep = false);
// the remote half of a copy-tag list: regions of boxes <-> one contiguous message ([tag][comp][cell]); to_buf = pack
// (source side of the tag), else unpack (destination side); skip_keep: values carrying the KEEP marker of
// k_patch_avgdown are not stored.  `base` = doubles already in the message (several arrays
// share one message), `cells_total` = cells of all tags of the message
int launch_patch_pack(const PBox* tab, int cur, const CopyTag* tags, int ntags, int arr, int ncomp, long long max_cells,
                      double* buf, long long base, bool to_buf, cudaStream_t st, bool skip_keep = false);
int launch_patch_fill(const PBox* tab, int nb, long long max_cells, int cur, int arr, int ncomp, double v, cudaStream_t st);
int launch_patch_initialize(const PBox* tab, int nb, long long max_cells, int cur, const BcInfo& B, const IcInfo& I,
                            cudaStream_t st);
int launch_patch_zero_solid(const PBox* tab, int nb, long long max_cells, int cur, cudaStream_t st);
int launch_patch_prepass(const PBox* tab, int nb, long long max_cells, int cur, const PGeom& G, cudaStream_t st);
// BCFill over faces, edges, corners of every box (26 launches at most; regions that are empty for every box are skipped
// by the caller through `any_outside`)
int launch_patch_physbc(const PBox* tab, int nb, long long max_face, int cur, const PGeom& G, const BcInfo& B,
                        cudaStream_t st);
int launch_patch_stream(const PBox* tab, int nb, long long max_cells, int cur, cudaStream_t st);
int launch_patch_qcorr(const PBox* tab, int nb, long long max_cells, int cur, const Phys& P, int want_macro, cudaStream_t st,
                       bool pull = false);
// fused stream + collide of a finest level: buffers [cur] -> [1 - cur]
int launch_patch_advance(const PBox* tab, int nb, long long max_cells, int cur, const PGeom& G, const Phys& P, int want_macro,
                         cudaStream_t st);
int launch_patch_collide(const PBox* tab, int nb, long long max_cells, int cur, const PGeom& G, const Phys& P, int want_macro,
                         cudaStream_t st);
// compute_eb_forces over the local boxes of a level: d_out3[3] (device) = the level's sum
int launch_patch_eb_forces(const PBox* tab, int nb, long long max_cells, int cur, double* d_out3, cudaStream_t st);
int launch_patch_derived(const PBox* tab, int nb, long long max_cells, const PGeom& G, const Phys& P, int with_dq,
                         cudaStream_t st);
// masked_avgdown (Utilities.H:315-350) of the fine boxes into the coarsened aux boxes (ng ghost cells), ratio 2
int launch_patch_avgdown(const PBox* ftab, int fcur, const PBox* ctab, int nb, long long max_cells, int ng, cudaStream_t st);
// CellConservativeLinear (mcslope, AMReX_MFInterp_3D_C.H:176-249) from the coarse aux boxes into fine ghost regions
int launch_patch_interp(const PBox* ftab, int fcur, const PBox* ctab, const RegionTag* regs, int nregs, long long max_cells,
                        cudaStream_t st);

}  // namespace mbl
