/* hifihr_b200 — C-ABI of the B200-native hand-render hot path.
 *
 * Plain pointers and sizes only (no torch types).  All pointers are DEVICE pointers
 * unless a field says "host".  Every entry point enqueues work on `stream`
 * (a cudaStream_t passed as void*), never allocates, never synchronises, and
 * returns 0 on success or an HFR_E* code; hfr_last_error() gives the message of
 * the calling thread's last failure.
 *
 * The reference has no FFI of its own (it is pure Python over PyTorch and
 * PyTorch3D).  Each entry point cites the reference call / upstream native
 * function it replaces; INTEGRATION.md shows the ctypes binding a maintainer adds.
 */
#ifndef HIFIHR_B200_H
#define HIFIHR_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HFR_OK 0
#define HFR_EINVAL 1    /* bad argument (shape, K out of range, null pointer)      */
#define HFR_ECUDA 2     /* CUDA runtime error at launch                            */
#define HFR_EUNSUPPORTED 3
#define HFR_ABI_VERSION 5

#define HFR_MAX_JOINTS 32
#define HFR_MAX_K 16

const char* hfr_last_error(void);
int hfr_abi_version(void);
/* 1 when a CUDA device of compute capability 10.x is current; else 0. */
int hfr_device_ok(void);

/* ------------------------------------------------------------------ articulated hand model
 * Constants of an LBS hand model (MANO: utils/my_mano.py:283-313 buffers; the
 * NIMBLE-shaped stand-in uses the same struct).  Built once by the host. */
typedef struct HfrHandModel {
  int32_t V;             /* vertices (778)                                               */
  int32_t NJ;            /* chain joints (16), joint 0 is the root                       */
  int32_t NS;            /* shape coefficients (10)                                      */
  int32_t NPC;           /* pose PCA coefficients consumed after the 3 root values (45); 0 = no PCA */
  int32_t NW;            /* skinning influences stored per vertex (<= 8)                 */
  int32_t NT;            /* tip vertices appended to the chain joints (5)                */
  int32_t center_joint;  /* index into the OUTPUT joint order to centre on (9), -1 = none */
  int32_t C3;            /* row pitch of `dirs` in floats (>= 3V, multiple of 4)          */
  const float* dirs;       /* (NS + 9(NJ-1), C3): shape then pose blend basis, coefficient-major */
  const float* v_template; /* (3V)                                                        */
  const float* J_template; /* (NJ,3)   = J_regressor @ v_template                         */
  const float* J_shapedirs;/* (NJ,3,NS)= J_regressor @ shapedirs                          */
  const float* pca_comps;  /* (NPC, 3(NJ-1)) selected components, or NULL                 */
  const float* pose_mean;  /* (3(NJ-1)) hands_mean, or NULL                              */
  const int32_t* parents;  /* (NJ) parent joint, -1 for the root                          */
  const int32_t* skin_idx; /* (NW, V) joint index per influence                           */
  const float* skin_w;     /* (NW, V) weight per influence (0 padded)                     */
  const int32_t* tip_verts;/* (NT)                                                        */
  const int32_t* joint_order; /* (NJ+NT) output joint k = chain∪tips[joint_order[k]]     */
  int32_t palm_verts[2];   /* root_palm mode (my_mano.py:459-461): the two vertices whose midpoint replaces
                              chain joint 0 in the joint output (MANO: 95, 22)            */
  /* optional joint-major view of the skinning weights (CSR over joints) for the backward's
   * d/d(joint transform) reduction; NULL = reduce vertex-major with warp sums */
  const int32_t* jv_ptr;   /* (NJ+1)                                                       */
  const int32_t* jv_vert;  /* (nnz) vertex of each non-zero weight                         */
  const float* jv_w;       /* (nnz)                                                        */
  /* optional: the blend basis pre-split into tf32 hi / lo parts and pre-tiled in the tensor-core operand layout
   * (hfr_mano_packed_basis_bytes / hfr_mano_pack_basis, built once).  With it, and a workspace in the call, the
   * layer runs as ONE batched blend contraction on the tensor cores (tcgen05) instead of one basis stream per
   * sample; NULL = per-sample kernels */
  const void* basis_packed;
} HfrHandModel;
/* Batched path (utils/my_mano.py:386-393 as one [B x NK].[NK x 3V] product per batch, and its transpose in the
 * backward): size of the packed basis, the one-time packing launch, and the per-call workspace size for B samples. */
int64_t hfr_mano_packed_basis_bytes(const HfrHandModel* m);
int hfr_mano_pack_basis(const HfrHandModel* m, void* packed, void* stream);
int64_t hfr_mano_workspace_bytes(const HfrHandModel* m, int32_t B);
/* 0 when no barrier wait of the batched kernels has ever timed out on the current device (diagnostic) */
int hfr_mano_batched_status(void);

/* Replaces ManoLayer.forward (utils/my_mano.py:315-483) / MyMANOLayer.forward (:39-54).
 * pose (B, 3+NPC) [or (B, 3*NJ) when NPC==0], betas (B,NS) or NULL (mean shape);
 * trans (B,3) or NULL (then centre on center_joint).  verts (B,V,3), joints (B,NJ+NT,3).
 * Non-default rotation inputs (my_mano.py:355-373): the first `n_rot_in` chain joints take their
 * rotation MATRIX from rots (B,n_rot_in,3,3) instead of Rodrigues — 1 = root_rot_mode 'rot6d' (the host
 * turns the 6-D vector into a matrix), NJ = joint_rot_mode 'rotmat' (pose may then be NULL).  `pose_off`
 * = number of leading pose columns that belong to the root (3 axis-angle, 6 rot6d); 0 means 3.
 * root_palm != 0: output joint sourced from chain joint 0 becomes the palm midpoint (:459-461). */
typedef struct HfrManoFwdArgs {
  int32_t B;
  const float* pose;
  const float* betas;
  const float* trans;
  float* verts;
  float* joints;
  const float* rots;
  int32_t n_rot_in, pose_off, root_palm;
  void* workspace;         /* hfr_mano_workspace_bytes(m, B) bytes, or NULL (per-sample kernels).  Selects the batched
                              tensor-core path when the model carries basis_packed.  Scratch of ONE call at a time:
                              calls that may run concurrently (different streams) need their own */
} HfrManoFwdArgs;
int hfr_mano_forward(const HfrHandModel* m, const HfrManoFwdArgs* a, void* stream);

/* Backward of the above (autograd of my_mano.py:315-483).  g_joints may be NULL.
 * Outputs g_pose (B,pose_off+NPC) (columns of a matrix-driven root are zeroed; may be NULL when
 * n_rot_in == NJ), g_betas (B,NS) (may be NULL), g_trans (B,3) (may be NULL). */
typedef struct HfrManoBwdArgs {
  int32_t B;
  const float* pose;
  const float* betas;
  const float* trans;
  const float* g_verts;
  const float* g_joints;
  float* g_pose;
  float* g_betas;
  float* g_trans;
  const float* rots;                /* as in the forward                                   */
  int32_t n_rot_in, pose_off, root_palm;
  float* g_rots;                    /* (B,n_rot_in,3,3) or NULL                            */
  void* workspace;                  /* as in the forward; NULL = per-sample kernel           */
  int32_t reuse_forward;            /* 1: `workspace` still holds what hfr_mano_forward left for the SAME inputs (pose
                                       state and posed rest vertices are not recomputed); 0: recompute them          */
} HfrManoBwdArgs;
int hfr_mano_backward(const HfrHandModel* m, const HfrManoBwdArgs* a, void* stream);

/* ------------------------------------------------------------------ per-sample geometry
 * Joint regression from POSED verts + FreiHAND reorder (xyz_from_vertice,
 * utils/Freihand_GNN_mano/Freihand_trainer_mano_fullsup.py:175-215), root shift
 * (models_res_nimble.py:159-166, 203-205), NDC projection (PerspectiveCameras + the
 * z overwrite of MeshRasterizer.transform; intrinsics from :228-235) and area-weighted
 * vertex normals (Meshes.verts_normals_packed).  Topology is shared by the batch. */
typedef struct HfrTopology {
  int32_t V, F;
  const int32_t* faces;       /* (F,3) vertex ids                                       */
  const int32_t* vf_ptr;      /* (V+1) CSR: incident (face*4+corner) entries per vertex  */
  const int32_t* vf_idx;      /* (3F)                                                    */
  /* sparse joint regressor on posed verts; NULL/0 when unused */
  int32_t NJR;                /* regressed joints (16)                                   */
  int32_t NOUT;               /* output joints (21)                                      */
  const int32_t* jr_ptr;      /* (NJR+1) CSR over joints                                 */
  const int32_t* jr_col;      /* (nnz) vertex ids                                        */
  const float* jr_val;        /* (nnz)                                                   */
  const int32_t* vj_ptr;      /* (V+1) CSC view for the backward                         */
  const int32_t* vj_row;      /* (nnz) joint ids                                         */
  const float* vj_val;        /* (nnz)                                                   */
  const int32_t* out_src;     /* (NOUT) >=0: regressed joint id; <0: -(vertex id)-1      */
  const int32_t* vf_nbr;      /* optional (3F,2): for incidence entry e of vertex a, the other two corners (b, c) of
                                 that face in winding order - saves the vf_idx -> faces round trip in the normal
                                 gathers; NULL = derive them from vf_idx / faces                                  */
} HfrTopology;

typedef struct HfrGeomFwdArgs {
  int32_t B;
  int32_t root_out;           /* output joint used as predicted root (9); -1 = no joints / no shift */
  const float* verts;         /* (B,V,3) hand-layer verts                                */
  const float* root_xyz;      /* (B,3) GT root in camera frame, or NULL                  */
  const float* focal;         /* (B,2) NDC focal as given to the camera (reference: -fcl) */
  const float* prp;           /* (B,2) NDC principal point                               */
  float* joints;              /* (B,NOUT,3) root-relative joints, or NULL                */
  float* verts_rel;           /* (B,V,3) verts - pred_root, or NULL                      */
  float* verts_view;          /* (B,V,3) verts - pred_root + root_xyz                    */
  float* verts_ndc;           /* (B,V,3) x_ndc, y_ndc, view z; or NULL                   */
  float* vnormals;            /* (B,V,3) or NULL                                         */
  float* face_verts;          /* (B,F,3,3) packed NDC face verts for the rasterizer, or NULL (needs verts_ndc) */
} HfrGeomFwdArgs;
int hfr_geom_forward(const HfrTopology* t, const HfrGeomFwdArgs* a, void* stream);

typedef struct HfrGeomBwdArgs {
  int32_t B;
  int32_t root_out;
  const float* verts;         /* forward input                                           */
  const float* root_xyz;
  const float* focal;
  const float* prp;
  const float* g_joints;      /* any of the g_* inputs may be NULL                       */
  const float* g_verts_rel;
  const float* g_verts_view;
  const float* g_verts_ndc;
  const float* g_vnormals;
  float* g_verts;             /* (B,V,3)                                                 */
  /* optional: gather d/d(view) and d/d(vertex normal) from the (face, tile) records of hfr_shade_backward_tiled instead
   * of reading g_verts_view / g_vnormals / g_verts_ndc (which must then be NULL).  Fixed summation order: incident
   * faces in CSR order, tiles of a face's range row by row. */
  const float* face_rec;
  const void* raster_ws;      /* the rasterizer workspace (tile ranges + record offsets), Ftot = B * F, mesh n = faces [n F, (n+1) F) */
  const uint32_t* status;     /* record-store status word of the backward; non-zero -> g_verts = NaN */
  /* optional scratch of hfr_geom_rec_partial_floats(t, B) floats: the records are first summed per (sample, incidence entry)
   * by a chip-wide launch (one thread per entry: the dependent range -> slot -> record loads of all entries overlap), the
   * per-sample kernel then adds a vertex's entries in CSR order (a fixed order: reproducible run to run) */
  float* rec_partial;
} HfrGeomBwdArgs;
int64_t hfr_geom_rec_partial_floats(const HfrTopology* t, int32_t B);
int hfr_geom_backward(const HfrTopology* t, const HfrGeomBwdArgs* a, void* stream);

/* Packed per-face vertices for the rasterizer: face_verts[b][f][c][:] = verts[b][faces[f][c]][:].
 * Replaces `verts_packed()[faces_packed()]` of pytorch3d MeshRasterizer.forward (reached from
 * models_res_nimble.py:208) and ATen's sort-based index backward behind it.  The backward adds a vertex's
 * incident (face, corner) entries in CSR order: no atomics, reproducible run to run. */
typedef struct HfrFaceVertsArgs {
  int32_t B;
  const float* verts;         /* forward: (B,V,3), e.g. verts_ndc                        */
  float* face_verts;          /* forward: (B,F,3,3)                                      */
  const float* g_face_verts;  /* backward: (B,F,3,3)                                     */
  float* g_verts;             /* backward: (B,V,3), overwritten                          */
} HfrFaceVertsArgs;
int hfr_face_verts_forward(const HfrTopology* t, const HfrFaceVertsArgs* a, void* stream);
int hfr_face_verts_backward(const HfrTopology* t, const HfrFaceVertsArgs* a, void* stream);

/* ------------------------------------------------------------------ rasterizer
 * Replaces pytorch3d._C.rasterize_meshes / rasterize_meshes_backward as reached from
 * MeshRasterizer.forward (models_res_nimble.py:208).  Packed face_verts (Ftot,3,3) in
 * NDC with view-space z; outputs are the four Fragments tensors, -1 filled.
 * K <= HFR_MAX_K.  `workspace` needs hfr_raster_workspace_bytes(Ftot) bytes. */
typedef struct HfrRasterArgs {
  int32_t N, H, W, K;
  int64_t Ftot;
  const float* face_verts;          /* (Ftot,3,3)                                         */
  const int64_t* mesh_first;        /* (N) first packed face of each mesh                 */
  const int64_t* mesh_nfaces;       /* (N)                                                */
  float blur_radius;
  int32_t perspective_correct, clip_barycentric, cull_backfaces;
  int64_t* pix_to_face;             /* (N,H,W,K)                                          */
  float* zbuf;                      /* (N,H,W,K)                                          */
  float* bary;                      /* (N,H,W,K,3)                                        */
  float* dists;                     /* (N,H,W,K)                                          */
  void* workspace;
  /* optional cost-ordered tile queue (hfr_raster_queue_bytes(N, H, W) bytes): the setup pass counts the faces whose tile
   * range covers each 16x16 tile and lists the tiles heaviest first; the fused kernel then runs them in that order with
   * the tiles no face touches (pure -1 fills, streamed with bulk shared->global copies) spread evenly in between, so
   * the ALU-bound and the HBM-bound work overlap and no heavy tile is left for the tail.  NULL = row-major tile order.
   * Used by hfr_raster_shade_forward. */
  void* tile_queue;
} HfrRasterArgs;
int64_t hfr_raster_workspace_bytes(int64_t Ftot);
int64_t hfr_raster_queue_bytes(int32_t N, int32_t H, int32_t W);
int hfr_raster_forward(const HfrRasterArgs* a, void* stream);

typedef struct HfrRasterBwdArgs {
  int32_t N, H, W, K;
  int64_t Ftot;
  const float* face_verts;
  const int64_t* pix_to_face;
  const float* g_zbuf;              /* (N,H,W,K)   may be NULL                            */
  const float* g_bary;              /* (N,H,W,K,3) may be NULL                            */
  const float* g_dists;             /* (N,H,W,K)   may be NULL                            */
  float blur_radius;
  int32_t perspective_correct, clip_barycentric;
  float* g_face_verts;              /* (Ftot,3,3), ACCUMULATED into (caller zeroes)       */
} HfrRasterBwdArgs;
int hfr_raster_backward(const HfrRasterBwdArgs* a, void* stream);

/* ------------------------------------------------------------------ shading + blending
 * Replaces HardPhongShader / SoftPhongShader / SoftSilhouetteShader forward (+backward):
 * TexturesUV.sample_textures, interpolate_face_attributes, phong_shading,
 * DirectionalLights, Materials and hard_rgb_blend / sigmoid_alpha_blend /
 * softmax_rgb_blend (reference construction models_res_nimble.py:79-96, 187-190). */
#define HFR_BLEND_HARD 0
#define HFR_BLEND_SIGMOID_ALPHA 1   /* SoftSilhouetteShader (colour = ones unless shading on) */
#define HFR_BLEND_SOFTMAX 2
#define HFR_SHADE_ONES 0            /* silhouette shader: colours are 1                   */
#define HFR_SHADE_PHONG_UV 1        /* Phong lighting x UV texture                        */

typedef struct HfrShadeParams {
  int32_t N, H, W, K;
  int32_t F, V;                     /* per-mesh faces / verts (shared topology)            */
  int32_t blend, shade;
  float sigma, gamma, znear, zfar;
  float background[3];
  float light_ambient[3], light_specular[3];
  float mat_ambient[3], mat_diffuse[3], mat_specular[3];
  float shininess;
  int32_t tex_n, tex_h, tex_w;      /* texture maps (tex_n = 1 shared, or N)               */
  int32_t VT;                       /* number of uv vertices                               */
  int32_t tex_pca;                  /* > 0: `texture` is the MEAN map (tex_n = 1) of a PCA texture model with this many
                                       components, evaluated per fragment: texel = mean + sum_k params[n][k] * basis[k]
                                       (NIMBLE-style per-sample texture without materialising the per-sample maps) */
  int32_t light_point;              /* 0: DirectionalLights, light_dir = direction (N,3).  1: PointLights (the branch of
                                       models_res_nimble.py:191-198 taken when ifLight=False): light_dir = LOCATION (N,3),
                                       the light direction of a fragment is location - position; g_light_dir then receives
                                       d/d(location) */
  int32_t tex_basis_stride;         /* layout of tex_basis.  0: component-major (tex_pca,tex_h,tex_w,3).  > 0: TEXEL-major
                                       (tex_h*tex_w, stride) floats, component k / channel c of a texel at 3k + c, zero
                                       padded; stride = 12 * ceil(tex_pca / 4) (16-byte aligned records): the tex_pca
                                       values a bilinear tap needs are one contiguous record instead of tex_pca reads
                                       tex_h*tex_w*12 bytes apart */
} HfrShadeParams;
#define HFR_MAX_TEX_PCA 64

typedef struct HfrShadeFwdArgs {
  HfrShadeParams p;
  const int64_t* pix_to_face; const float* zbuf; const float* bary; const float* dists;
  const int32_t* faces;             /* (F,3)                                               */
  const float* verts_view;          /* (N,V,3)                                             */
  const float* vnormals;            /* (N,V,3)                                             */
  const int32_t* faces_uvs;         /* (F,3)                                               */
  const float* verts_uvs;           /* (VT,2)                                              */
  const float* texture;             /* (tex_n,tex_h,tex_w,3)                               */
  const float* light_dir;           /* (N,3)                                               */
  const float* light_color;         /* (N,3) diffuse colour                                */
  float* image;                     /* (N,H,W,4)                                           */
  /* optional: per-(mesh, face) attribute records written by hfr_face_attr_forward, (N,F,HFR_FACE_ATTR_FLOATS).
   * When set, the shaders read one contiguous record per fragment instead of chasing
   * faces -> verts_view / vnormals and faces_uvs -> verts_uvs (two dependent gathers).  NULL = gather path. */
  const float* face_attr;
  const float* tex_basis;           /* when p.tex_pca > 0; layout: p.tex_basis_stride      */
  const float* tex_params;          /* (N,tex_pca)             when p.tex_pca > 0          */
} HfrShadeFwdArgs;
int hfr_shade_forward(const HfrShadeFwdArgs* a, void* stream);

/* Per-face attribute records for the shaders (what interpolate_face_attributes / TexturesUV.sample_textures
 * gather per fragment upstream): for every mesh n and face f, 28 floats =
 *   view-space corner positions (9), corner vertex normals (9), corner uvs (6), corner vertex ids (3, int bits), pad. */
#define HFR_FACE_ATTR_FLOATS 28
typedef struct HfrFaceAttrArgs {
  int32_t N, F, V, VT;
  const int32_t* faces;             /* (F,3)                                               */
  const float* verts_view;          /* (N,V,3)                                             */
  const float* vnormals;            /* (N,V,3)                                             */
  const int32_t* faces_uvs;         /* (F,3)                                               */
  const float* verts_uvs;           /* (VT,2)                                              */
  float* face_attr;                 /* (N,F,28)                                            */
} HfrFaceAttrArgs;
int hfr_face_attr_forward(const HfrFaceAttrArgs* a, void* stream);

typedef struct HfrShadeBwdArgs {
  HfrShadeFwdArgs f;                /* forward inputs; f.image = the forward OUTPUT (read by the softmax blend) */
  const float* g_image;             /* (N,H,W,4)                                           */
  /* dense per-fragment grads for the modular (autograd) path; any may be NULL */
  float* g_zbuf; float* g_bary; float* g_dists;
  /* fused path: when g_verts_ndc != NULL the rasterizer backward is applied in the same
   * kernel and accumulated per vertex (needs face_verts_ndc = verts_ndc, (N,V,3)). */
  const float* verts_ndc; float* g_verts_ndc;
  float blur_radius; int32_t perspective_correct, clip_barycentric;
  /* accumulated (caller zeroes) */
  float* g_verts_view;              /* (N,V,3)                                             */
  float* g_vnormals;                /* (N,V,3)                                             */
  float* g_texture;                 /* (tex_n,tex_h,tex_w,3)                               */
  float* g_light_dir;               /* (N,3)                                               */
  float* g_light_color;             /* (N,3)                                               */
  /* optional: per-mesh tile box the rasterizer's setup pass left in its workspace (hfr_raster_tile_box);
   * tiles outside a mesh's footprint then leave without reading the Fragments.  NULL disables. */
  const uint32_t* tile_box;
  /* SSAA-fused path (hfr_raster_shade_pool_forward): when pool_aa > 1, g_image is the gradient of the POOLED
   * RGBA image, (N,H/pool_aa,W/pool_aa,4); every rasterised pixel takes g/pool_aa^2 of its pooled pixel
   * (avg_pool2d backward, models_res_nimble.py:211) and, with pool_binarize, no gradient reaches alpha (the
   * reference binarises re_sil in place, :219).  0 or 1 = g_image is full resolution. */
  int32_t pool_aa, pool_binarize;
  float* g_tex_params;              /* (N,tex_pca) accumulated (caller zeroes), PCA textures only; may be NULL */
} HfrShadeBwdArgs;
int hfr_shade_backward(const HfrShadeBwdArgs* a, void* stream);
/* Address of the (N,4) uint32 tile box {txmin, 255-txmax, tymin, 255-tymax} (16x16-pixel tiles) inside a
 * rasterizer workspace that hfr_raster_forward / hfr_raster_shade_forward filled for N meshes and Ftot packed
 * faces; NULL when the workspace holds none (N > Ftot). */
const uint32_t* hfr_raster_tile_box(const void* workspace, int64_t Ftot, int32_t N);

/* Atomics-free, deterministic fused backward (shade' + blend' + rasterize' + d(ndc)/d(view)) - the replacement of
 * upstream rasterize_meshes_backward / interp_face_attrs_backward's 9 + 18 atomics per covered pixel, reached by
 * loss.backward() at train_hrnet.py:112.  One CTA owns a 16x16 tile: its fragments are counting-sorted by face in
 * shared memory (integer bookkeeping only), each face's 18 gradient components (3 corners x {view position, vertex
 * normal}) are summed SEQUENTIALLY in pixel order by one warp and stored as ONE record per (face, tile) at the slot the
 * rasterizer's setup pass reserved (hfr_raster_workspace layout); hfr_geom_backward then gathers per vertex over the
 * static vertex->face incidence lists in a fixed order.  Quantities summed over ALL fragments of a sample or batch
 * (light gradients, the shared texture's gradient) go through 64-bit FIXED-POINT accumulators: integer addition is
 * associative, so the result does not depend on the order the hardware serialises them in; hfr_grad_finish converts
 * them to fp32 and clears them.  The whole backward is therefore bit-reproducible run to run.
 * Needs the rasterizer workspace exactly as hfr_raster_forward or hfr_raster_shade_forward left it for the same N / Ftot, with
 * mesh_first[n] = n * F (uniform packing, what MeshRasterizer builds for a batch of one topology). */
#define HFR_FACE_REC_FLOATS 18      /* corner-major: {d/d(view xyz), d/d(vertex normal xyz)} x 3 corners */
typedef struct HfrShadeBwdTiledArgs {
  HfrShadeFwdArgs f;                /* forward inputs; f.image = the forward OUTPUT (softmax blend)             */
  const float* g_image;             /* (N,H,W,4), or the pooled gradient (N,H/aa,W/aa,4) when pool_aa > 1          */
  const float* verts_ndc;           /* (N,V,3)                                                                     */
  const float* focal;               /* (N,2) NDC focal as given to the camera: d(ndc)/d(view) is folded in here    */
  float blur_radius; int32_t perspective_correct, clip_barycentric;
  const void* raster_ws;            /* rasterizer workspace filled by the forward                                  */
  float* face_rec;                  /* (rec_cap, HFR_FACE_REC_FLOATS) record store; cleared here up to the used size */
  int64_t rec_cap;                  /* records that fit; when the batch needs more: status bit 0, gradients = NaN  */
  int64_t* light_acc;               /* (N,6) fixed point d/d(light_dir | location), d/d(light_color); caller zeroes ONCE,
                                       hfr_grad_finish re-clears                                                    */
  int64_t* tex_acc;                 /* (tex_n,tex_h,tex_w,3) fixed-point gradient of the texture, same protocol; NULL:
                                       plain fp32 atomics into g_texture (faster, not run-to-run reproducible)      */
  float* g_texture;                 /* used when tex_acc == NULL (accumulated, caller zeroes); may be NULL           */
  float* fx_scale;                  /* DEVICE float: the fixed-point multiplier (a power of two).  READ when gmax_bits is NULL
                                       (pick 2^34 / the power of two above max |g_image|); WRITTEN when gmax_bits is set */
  uint32_t* status;                 /* DEVICE word, OR-ed: bit 0 = record store too small                            */
  int32_t pool_aa, pool_binarize;   /* as in HfrShadeBwdArgs                                                         */
  const uint32_t* gmax_bits;        /* optional: bit pattern of max |g_image| as hfr_loss_backward leaves it; the multiplier
                                       is derived from it on the device (no host round trip, no extra reduction pass)   */
  /* optional mean-RGB fix-up (pairs with HfrLossBwdArgs.skip_mrgb): g_image lacks the gradient of the mean-RGB term
   * (losses.py:369), which needs the GLOBAL sums; it is added here per pixel from fix_sums[HFR_LOSS_SUM_R / _T]
   * (all-reduced), fix_w[1] = d(total)/d(mrgb), the image the loss saw (fix_image, (N,H,W,4) or pooled) and
   * fix_inv_scale = 1 / sil_scale, fix_count = N_global * 3 * H * W of the loss resolution */
  const float* fix_sums; const float* fix_w; const float* fix_image;
  float fix_inv_scale; int64_t fix_count;
  /* optional: the tile queue the forward filled (HfrRasterArgs.tile_queue, same N / H / W): only tiles that hold faces are
   * visited, heaviest class first.  NULL = every tile of the grid, row-major */
  const void* tile_queue;
} HfrShadeBwdTiledArgs;
int hfr_shade_backward_tiled(const HfrShadeBwdTiledArgs* a, void* stream);

/* Fixed-point accumulators -> fp32 gradients (and clears the accumulators for the next step):
 *   g_texture[i] = tex_acc[i] / scale (n_tex values), g_light_dir / g_light_color (N,3) from light_acc (N,6). */
typedef struct HfrGradFinishArgs {
  int64_t* tex_acc; float* g_texture; int64_t n_tex;
  int64_t* light_acc; float* g_light_dir; float* g_light_color; int32_t N;
  const float* fx_scale;
  uint32_t* gmax_bits;              /* optional: reset to 0 for the next step                                         */
} HfrGradFinishArgs;
int hfr_grad_finish(const HfrGradFinishArgs* a, void* stream);

/* Fused rasterize + shade forward: writes Fragments AND the image in one pass. */
typedef struct HfrRasterShadeArgs {
  HfrRasterArgs r;
  HfrShadeFwdArgs s;                /* its Fragments pointers are ignored (taken from r)   */
} HfrRasterShadeArgs;
int hfr_raster_shade_forward(const HfrRasterShadeArgs* a, void* stream);

/* Fused rasterize + shade + SSAA pool + output split: the reference's own render setting
 * (models_res_nimble.py:74-96 image_size=672, faces_per_pixel=1; :208-220 render, avg_pool2d(3,3),
 * RGB/alpha split, in-place binarisation, maskRGBs) in ONE pass: Fragments are written at the
 * rasterised resolution (r.H, r.W), the full-resolution RGBA image is never materialised
 * (s.image may be NULL) and the pooled outputs leave the SM directly.
 * r.H and r.W must be multiples of aa.  pooled is required; the NCHW outputs are optional. */
typedef struct HfrRasterShadePoolArgs {
  HfrRasterArgs r;
  HfrShadeFwdArgs s;                /* Fragments pointers ignored (taken from r); image optional */
  int32_t aa, binarize;
  const float* images_in;           /* (N,3,H/aa,W/aa) network input for mask_rgbs, or NULL */
  float* pooled;                    /* (N,H/aa,W/aa,4) RGBA, alpha binarised when `binarize` (loss kernels, nhwc=1) */
  float* re_img;                    /* (N,3,H/aa,W/aa) or NULL                              */
  float* re_sil;                    /* (N,1,H/aa,W/aa) or NULL                              */
  float* mask_rgbs;                 /* (N,3,H/aa,W/aa) or NULL (needs images_in)            */
} HfrRasterShadePoolArgs;
int hfr_raster_shade_pool_forward(const HfrRasterShadePoolArgs* a, void* stream);

/* ------------------------------------------------------------------ SSAA pooling + output split
 * models_res_nimble.py:210-220: NHWC->NCHW, avg_pool2d(aa,aa), split RGB / alpha,
 * optional in-place binarisation of alpha>0 to 255, maskRGBs = images*(re_sil>0). */
typedef struct HfrPoolArgs {
  int32_t N, H, W, aa;              /* H,W = OUTPUT size; input is (N,H*aa,W*aa,4)          */
  int32_t binarize;
  const float* image;               /* (N,H*aa,W*aa,4)                                     */
  const float* images_in;           /* (N,3,H,W) network input for maskRGBs, or NULL       */
  float* re_img;                    /* (N,3,H,W)                                           */
  float* re_sil;                    /* (N,1,H,W)                                           */
  float* mask_rgbs;                 /* (N,3,H,W) or NULL                                   */
} HfrPoolArgs;
int hfr_pool_forward(const HfrPoolArgs* a, void* stream);
typedef struct HfrPoolBwdArgs {
  int32_t N, H, W, aa;
  int32_t binarize;                 /* when set, no gradient reaches alpha (reference mode) */
  const float* g_re_img;            /* (N,3,H,W) or NULL                                   */
  const float* g_re_sil;            /* (N,1,H,W) or NULL                                   */
  float* g_image;                   /* (N,H*aa,W*aa,4)                                     */
} HfrPoolBwdArgs;
int hfr_pool_backward(const HfrPoolBwdArgs* a, void* stream);

/* ------------------------------------------------------------------ losses
 * losses.py:355-378 (texture, mrgb, ssim_tex), :399-408 (sil, iou) with
 * utils/losses_util.py:366-378 and utils/pytorch_ssim/__init__.py:17-37.
 * Two phases so the global means of `mrgb` are known before gradients are formed:
 *   hfr_loss_forward : partial sums -> sums[HFR_LOSS_NSUMS + 2N] (caller zeroes), plus the
 *                      SSIM derivative maps needed by the backward (dmaps, 9 floats/pixel).
 *   hfr_loss_backward: reads sums (optionally all-reduced across ranks) -> g_re_img, g_re_sil.
 * Layout: re_img (N,3,H,W), re_sil (N,1,H,W), imgs (N,3,H,W), seg (N,H,W) float. */
#define HFR_LOSS_L1 0        /* sum |rim - target|            */
#define HFR_LOSS_SUM_R 1     /* sum rim                       */
#define HFR_LOSS_SUM_T 2     /* sum target                    */
#define HFR_LOSS_SIL 3       /* sum |re_sil - seg|            */
#define HFR_LOSS_SSIM 4      /* sum ssim_map                  */
#define HFR_LOSS_L2 5        /* sum (rim - target)^2, metric modes only (mask_mode != 0) */
#define HFR_LOSS_NSUMS 8     /* then per-sample: mul[N], add[N] for IoU */
typedef struct HfrLossArgs {
  int32_t N, H, W;
  float sil_scale;                  /* 255 (reference, binarised) or 1 (soft alpha)        */
  int32_t want_ssim, want_grad;
  int32_t nhwc;                     /* 1: re_img points at the shader's RGBA image (N,H,W,4), re_sil is ignored,
                                       and the backward writes g_re_img as (N,H,W,4) (fused path, no pooling) */
  const float* re_img; const float* re_sil; const float* imgs; const float* seg;
  float* sums;                      /* (HFR_LOSS_NSUMS + 2N)                               */
  const float* gauss;               /* DEVICE pointer to the 11 fp32 Gaussian taps, or NULL when !want_ssim */
  float* dmaps;                     /* (N,9,H,W) or NULL when !want_ssim || !want_grad      */
  uint8_t* tile_flags;              /* optional (N, ceil(H/4), ceil(W/4)): written by the forward (1 = that 4x4 pixel
                                       block holds a non-zero masked-image sample), read by the backward to skip the
                                       SSIM stencil where it contributes exactly nothing; NULL disables the skip */
  int32_t mask_mode;                /* 0: training losses (rim = re_img * re_sil / sil_scale, target = imgs * seg).
                                       Evaluation-time texture metrics (train_hrnet.py:149-161; forward only):
                                       1: both images * seg; 2: both images * (re_sil > 0) (the HO3D branch);
                                       sums[HFR_LOSS_L1], [HFR_LOSS_L2], [HFR_LOSS_SSIM] then give L1 / L2 / PSNR / SSIM
                                       3: the self-supervised photometric terms texture_self / mrgb_self / ssim_tex_self
                                       (losses.py:317-340): x = re_img as rendered (no silhouette factor), y = `imgs`, which
                                       the caller points at maskRGBs (models_res_nimble.py:220); re_sil / seg are unused
                                       (NCHW only).  sums needs HFR_LOSS_NSUMS + 3N floats: [HFR_LOSS_SSIM] the SSIM sum and,
                                       per sample n, [NSUMS + n] = sum|x - y|, [NSUMS + N + n] = sum x, [NSUMS + 2N + n] = sum y
                                       (the reference weights them by texture_con[n]^2).  Has a backward. */
  /* optional 8-bit transport of the targets (the datasets store 8-bit images and {0,1} masks; the reference
   * converts on the host, ToTensor = x / 255): when set they REPLACE imgs / seg, and the kernels convert while
   * loading (exact x / 255.0f through a 256-entry table, mask byte != 0 -> 1.0f), so 4x fewer bytes cross PCIe. */
  const uint8_t* imgs_u8;           /* (N,3,H,W) or NULL                                   */
  const uint8_t* seg_u8;            /* (N,H,W) or NULL                                     */
  /* optional deterministic reduction: every CTA leaves its 8 partial sums in `partials` (hfr_loss_partials_floats(N,H,W)
   * floats, 16-byte aligned), the last CTA to arrive adds them in a fixed order and WRITES sums (the caller need not
   * zero them).  `ticket` = one DEVICE word, zero before the first launch (the kernel resets it).  NULL: one fp32
   * atomic per CTA and component (the sums then differ in the last bits from run to run). */
  float* partials;
  uint32_t* ticket;
  /* optional (fused step): the rasterizer's per-mesh tile box (hfr_raster_tile_box) and the ratio rasterised / loss
   * resolution.  The backward that is limited to that box (HfrLossBwdArgs.tile_box) reads the derivative maps only
   * within 5 pixels of the 32x32 tiles touching it, so the forward writes `dmaps` only there.  NULL = everywhere. */
  const uint32_t* dmaps_box;
  int32_t dmaps_box_aa;
} HfrLossArgs;
int64_t hfr_loss_partials_floats(int32_t N, int32_t H, int32_t W);
int hfr_loss_forward(const HfrLossArgs* a, void* stream);
typedef struct HfrLossBwdArgs {
  HfrLossArgs f;
  /* DEVICE pointer to 5 floats: d(total)/d(term) for texture, mrgb, ssim_tex, sil, iou
   * (lambda x upstream grad; kept on the device so the backward needs no host sync) */
  const float* w;
  const float* gauss;               /* DEVICE pointer to the 11 fp32 Gaussian taps (sigma 1.5) */
  int64_t count_global;             /* N_global*3*H*W for the means (multi-GPU aware)      */
  int32_t n_global;                 /* global batch for the IoU mean                       */
  float* g_re_img; float* g_re_sil; /* (N,3,H,W), (N,1,H,W)                                */
  /* mask_mode 3 only: w[0..2] = d(total)/d(texture_self, mrgb_self, ssim_tex_self); g_re_sil is not written */
  const float* tex_con;             /* (N) examples['texture_con'] (losses.py:325)         */
  const float* self_norm;           /* DEVICE pointer to 1 float: sum_n tex_con[n]^2 over the GLOBAL batch */
  /* fused step only (all optional) */
  const uint32_t* tile_box;         /* per-mesh tile box of the rasterizer (hfr_raster_tile_box): tiles whose pixels hold no
                                       fragment are not written (their gradient is never read by the fused backward)  */
  int32_t box_aa;                   /* rasterised resolution / loss resolution (SSAA factor), 0 or 1 = same           */
  int32_t skip_mrgb;                /* 1: leave the mean-RGB term out; hfr_shade_backward_tiled adds it from the sums  */
  uint32_t* gmax_bits;              /* DEVICE word, atomicMax of the bit pattern of max |gradient written| (caller zeroes
                                       or lets hfr_grad_finish reset it): scales the fixed-point accumulators           */
} HfrLossBwdArgs;
int hfr_loss_backward(const HfrLossBwdArgs* a, void* stream);

/* ------------------------------------------------------------------ keypoints + mesh regularisers
 * SURVEY.md §8(f) rows 2-3.  j2d = proj_func(joints + root_xyz, K) (utils/traineval_util.py:338-354 as called at
 * train_hrnet.py:83; utils/fh_utils.py:30-39) and the keypoint / mesh terms of LossFunction.__call__
 * (losses.py:244-299): joint_2d, joint_3d, vert_3d (base_loss_fn = L1 mean or MSE), bone_direc, bone_direc_3d
 * (utils/losses_util.py:217-282), edge_length (:284-301), mscale.  One launch forward, one backward.
 *   forward : writes j2d (optional) and ADDS per-term partial sums to sums[HFR_KP_NSUMS] (caller zeroes; under
 *             data parallelism the caller all-reduces them);  term = sums[k] / count_k with counts
 *             n*NJ*2, n*NJ*3, n*V*3, n*NB, n*NB, n*3F, n, n*V  (n = global batch).
 *   backward: w[k] = d(total)/d(term k) (DEVICE floats, lambda x upstream) -> g_joints (B,NJ,3) and g_verts (B,V,3),
 *             both WRITTEN (not accumulated); feed them to hfr_geom_backward as g_joints / g_verts_rel. */
#define HFR_KP_J2D 0
#define HFR_KP_J3D 1
#define HFR_KP_V3D 2
#define HFR_KP_BONE2D 3
#define HFR_KP_BONE3D 4
#define HFR_KP_EDGE 5
#define HFR_KP_MSCALE 6
#define HFR_KP_LAP 7      /* 'triangle': uniform Laplacian smoothing (losses.py:422-429, utils/losses_util.py:340-364); count n*V */
#define HFR_KP_NSUMS 8
typedef struct HfrKeypointArgs {
  int32_t B, NJ, V, F;
  int32_t l2;                       /* base_loss_fn: 0 = nn.L1Loss, 1 = mse_loss (losses.py:239-242)        */
  int32_t NB;                       /* bones                                                               */
  int32_t scale_a, scale_b;         /* mscale reference bone (9, 10); scale_a < 0 disables the term         */
  float scale_len;                  /* 0.0282                                                              */
  const float* joints;              /* (B,NJ,3) root-relative joints (hfr_geom_forward output)             */
  const float* root_xyz;            /* (B,3) or NULL                                                       */
  const float* Kmat;                /* (B,3,3) intrinsics, or NULL (no projection, no 2-D terms)           */
  const float* verts;               /* (B,V,3) root-relative verts, or NULL                                */
  const int32_t* faces;             /* (F,3)                                                               */
  const float* joints_gt;           /* (B,NJ,3) or NULL: joint_3d, bone_direc_3d off                       */
  const float* j2d_gt;              /* (B,NJ,2) or NULL: joint_2d, bone_direc off                          */
  const float* verts_gt;            /* (B,V,3)  or NULL: vert_3d, edge_length off                          */
  const float* conf;                /* (B,NJ) joint confidences for the bone terms, or NULL (= ones)       */
  const int32_t* bone_parent;       /* (NB)                                                                */
  const int32_t* bone_child;        /* (NB)                                                                */
  float* j2d;                       /* (B,NJ,2) output, or NULL                                            */
  float* sums;                      /* (HFR_KP_NSUMS) accumulated                                          */
  /* uniform Laplacian term: CSR of each vertex's neighbours over the mesh's unique edges; NULL disables it */
  const int32_t* nbr_ptr;           /* (V+1)                                                               */
  const int32_t* nbr_idx;           /* (2E)                                                                */
} HfrKeypointArgs;
int hfr_keypoint_forward(const HfrKeypointArgs* a, void* stream);
typedef struct HfrKeypointBwdArgs {
  HfrKeypointArgs f;
  const float* w;                   /* DEVICE (HFR_KP_NSUMS) d(total)/d(term)                              */
  int32_t n_global;                 /* global batch of the means                                           */
  const float* g_j2d_in;            /* (B,NJ,2) extra upstream gradient on the j2d output, or NULL         */
  float* g_joints;                  /* (B,NJ,3)                                                            */
  float* g_verts;                   /* (B,V,3) or NULL                                                     */
} HfrKeypointBwdArgs;
int hfr_keypoint_backward(const HfrKeypointBwdArgs* a, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HIFIHR_B200_H */
