/*
 * ohao_b200.h — C ABI of the B200-native OHAO path-tracing hot path.
 *
 * This is the drop-in boundary.  The reference has no C plugin
 * ABI; the seam this library sits behind is the C++ interface
 * IRTRendererProfile (ohao/render/rt/rt_profile_renderer.hpp:7-86) together
 * with RTAccelerationStructure (ohao/render/rt/rt_acceleration_structure.hpp:49-117)
 * and the scene->GPU packers of VulkanRenderer (ohao/gpu/vulkan/rt_build.cpp,
 * light_upload.cpp).  Every entry point below cites the reference interface it
 * replaces.  Plain pointers and sizes only; no torch / CUDA types.
 *
 * Conventions (mirroring the reference's "[[nodiscard]] bool + std::cerr"):
 *   - every int-returning call returns 0 on success, non-zero on failure, and
 *     ohb_last_error(ctx) then holds a one-line description;
 *   - one context per GPU, calls on one context are single-threaded;
 *   - all pointers are HOST pointers unless the name says "_dev";
 *   - matrices are column-major float[16] exactly as glm::mat4 stores them.
 *
 * There is no CPU fallback: ohb_create fails if no CUDA device is usable.
 */
#ifndef OHAO_B200_H
#define OHAO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OHB_ABI_VERSION 1u
#define OHB_NO_TEXTURE 0xFFFFFFFFu
#define OHB_MISS 0xFFFFFFFFu

typedef struct ohb_ctx ohb_ctx;

/* RTRenderProfile (rt_settings.hpp:11-14) */
enum { OHB_PROFILE_OFFLINE = 0, OHB_PROFILE_REALTIME = 1 };
/* SamplerType (sampler_types.hpp) ; GLSL SAMPLER_PCG=0 / SAMPLER_SOBOL=1 (sampler_api.glsl:12-13) */
enum { OHB_SAMPLER_PCG = 0, OHB_SAMPLER_SOBOL = 1 };
/* DenoiseMode (denoise/denoise_types.hpp:13-19), same numeric values.  Implemented: NONE and ATROUS (the SVGF
 * denoiser of atrous_denoise.cpp, realtime profile only); ohb_set_settings refuses the others (OIDN / NRD /
 * DLSS-RR are third-party libraries). */
enum { OHB_DENOISE_NONE = 0, OHB_DENOISE_OIDN = 1, OHB_DENOISE_NRD = 3, OHB_DENOISE_ATROUS = 4, OHB_DENOISE_DLSSRR = 5 };

/* Control-flag bits, identical to the raygen's PT_FLAG_* (pt_raygen_offline.rgen:59-63,
 * pt_raygen_realtime.rgen flag block) so a push-constant dump can be replayed. */
#define OHB_FLAG_ENABLE_AOVS             (1u << 0)
#define OHB_FLAG_ENABLE_INTERNAL_DENOISE (1u << 1)
#define OHB_FLAG_ENABLE_FIREFLY_CLAMP    (1u << 2)
#define OHB_FLAG_RESTIRGI_OFF            (1u << 5)   /* no temporal / spatial reuse (M = 1 anchor)          */
#define OHB_FLAG_RESTIRGI_GIONLY         (1u << 6)   /* measurement bit of the reference; affects NRD AOVs only */
#define OHB_FLAG_RESTIRGI_LEGACY         (1u << 7)   /* reference's multi-bounce Stage C escape hatch: NOT implemented, ohb_render refuses it */
#define OHB_FLAG_RESTIRGI_NOSPATIAL      (1u << 8)   /* temporal-only                                       */
/* Not a reference flag: reproduces the integrator revision that rendered the reference's
 * committed golden image tests/golden/cornell_box.png (Stage C spec-lobe continuation uses
 * mix(1,albedo,metallic) like Stage B instead of HEAD's albedo*(1-metallic),
 * pt_raygen_offline.rgen:864 vs :1172).  Off by default == HEAD behaviour. */
#define OHB_FLAG_GOLDEN_COMPAT           (1u << 16)

/* One TLAS instance — RTInstance (rt_acceleration_structure.hpp:38-44) as filled by
 * VulkanRenderer::buildBLASTLAS (rt_build.cpp:851-897): one BLAS per actor over a
 * slice of the shared index buffer, customIndex = running global triangle offset.
 * xform is the 3x4 ROW-major object->world matrix (VkTransformMatrixKHR layout,
 * i.e. the transpose of glm's upper 3 rows, rt_acceleration_structure.cpp:430-441). */
typedef struct ohb_instance {
    uint32_t first_tri;   /* == customIndex: first global triangle of this BLAS   */
    uint32_t tri_count;   /* triangles in the BLAS (indexCount / 3)               */
    float    xform[12];   /* row-major 3x4 object->world                          */
    uint32_t mask;        /* visibility mask, 0xFF in the reference               */
    uint32_t _pad;
} ohb_instance;

/* RTRenderSettings (rt_settings.hpp:17-35) — the fields the PT hot path reads. */
typedef struct ohb_settings {
    uint32_t profile;             /* OHB_PROFILE_*                                   */
    uint32_t max_bounces;         /* 4 offline / 2 realtime                          */
    uint32_t flags;               /* OHB_FLAG_* (AOVs / internal denoise / firefly)  */
    float    firefly_clamp_lum;   /* pc.tuning.x                                     */
    uint32_t sampler_type;        /* OHB_SAMPLER_* (both profiles run Sobol, Q3)     */
    float    anisotropy_strength; /* pc.jitter.z                                     */
    float    anisotropy_rotation; /* pc.jitter.w                                     */
    float    subsurface_strength; /* pc.tuning.w                                     */
    uint32_t samples_per_frame;   /* realtime N-spp loop, clamped to [1,64]          */
    uint32_t denoise_mode;        /* OHB_DENOISE_* (RTRenderSettings::denoiseMode)   */
    uint32_t _pad[2];
} ohb_settings;

/* Ray / hit records of the parity hook (north_star: closest-hit primitive IDs
 * bit-exact on recorded ray batches).  Interval is tmin < t < tmax as for
 * traceRayEXT triangle candidates (pt_raygen_offline.rgen:198-199). */
typedef struct ohb_ray { float origin[3]; float tmin; float dir[3]; float tmax; } ohb_ray;
typedef struct ohb_hit { float t; float u; float v; uint32_t prim; } ohb_hit; /* prim = global triangle id, OHB_MISS on miss; (u,v) = weights of vertex 1,2 */

/* Device counters (rays/sample and hits/sample are measured, never assumed). */
typedef struct ohb_counters {
    uint64_t samples;        /* path trees started                                  */
    uint64_t closest_rays;   /* closest-hit queries traced                          */
    uint64_t shadow_rays;    /* visibility queries traced                           */
    uint64_t closest_hits;   /* closest-hit queries that hit a triangle             */
    uint64_t kernel_launches;/* kernels launched by this context since last reset   */
    uint64_t _reserved[3];
} ohb_counters;

/* BVH statistics for DESIGN.md / bench (node count, SAH cost, build time). */
typedef struct ohb_accel_stats {
    uint32_t num_tris;       /* triangles the BVH was built over                    */
    uint32_t num_nodes;      /* 8-wide nodes (80 B each)                            */
    uint32_t levels;         /* depth of the 8-wide tree                            */
    uint32_t max_leaf_tris;  /* triangles per leaf child                            */
    float    sah_cost;       /* SAH cost of the final tree (Ct=1, Ci=1)             */
    float    build_ms;       /* device time of the last ohb_build_accel             */
    uint32_t treelet_passes;
    float    update_ms;      /* device time of the last ohb_update_instances (MODE_UPDATE refit)   */
} ohb_accel_stats;

/* ---- lifecycle: PathTracer::{init,destroy,resize} (path_tracer.cpp:73-184,213-260) ---- */
uint32_t    ohb_abi_version(void);
ohb_ctx*    ohb_create(int device_ordinal, uint32_t width, uint32_t height, int profile);
void        ohb_destroy(ohb_ctx*);
int         ohb_resize(ohb_ctx*, uint32_t width, uint32_t height);
const char* ohb_last_error(const ohb_ctx*);   /* "" when no error; valid for ctx==NULL after a failed create */

/* ---- scene data contract -------------------------------------------------- */
/* createRTVertexIndexBuffers / createRTNormalUVBuffers (rt_build.cpp:27-184) + matIDs
 * (rt_build.cpp:188-218): shared vertex buffer (position at byte 0 of each
 * stride_bytes record; the reference's Vertex is 100 B), GLOBAL uint32 indices,
 * vec4 normals, vec2 uvs, one material id per triangle. */
int ohb_set_geometry(ohb_ctx*, const void* positions, size_t stride_bytes, uint32_t nverts,
                     const uint32_t* global_indices, uint32_t ntris,
                     const float* normals_vec4, const float* uvs_vec2, const uint32_t* mat_ids);
/* RTAccelerationStructure::{clearInstances,addInstance} (rt_acceleration_structure.cpp:419-470) */
int ohb_set_instances(ohb_ctx*, const ohb_instance* instances, uint32_t n);
/* uploadRTMaterialBuffers / updateRTMaterialParams (rt_build.cpp:188-389): 3 vec4 per material */
int ohb_set_materials(ohb_ctx*, const float* mat_colors_3vec4, uint32_t nmaterials);
/* uploadRTTextureArray (rt_build.cpp:391-823): R8G8B8A8_UNORM layers of one size, LINEAR/REPEAT */
int ohb_set_textures(ohb_ctx*, const uint8_t* rgba8_layers, uint32_t w, uint32_t h, uint32_t layers);
/* uploadLightBuffer / updateRTLightParams (light_upload.cpp:145-293): 16-B header
 * {u32 lightCount, u32 envMapTexIdx, f32 envIntensity, u32 pad} + 80-B GPULight records */
int ohb_set_lights(ohb_ctx*, const void* light_ssbo, size_t bytes);
/* setEnvironmentMap + EnvCDF::build (light_upload.cpp:306-465, env_cdf.cpp:13-61): uploads the
 * RGBA32F equirect image and builds marginal/conditional CDFs ON THE DEVICE in the
 * reference's summation order.  rgba32f==NULL removes the env map (dummy 1-float CDFs, 0x0). */
int ohb_set_env(ohb_ctx*, const float* rgba32f, uint32_t w, uint32_t h);
/* Parity hook for north_star check (ii): copies the device-built CDFs back. */
int ohb_get_env_cdf(ohb_ctx*, float* marginal_h, float* conditional_wh, float* integral);
/* Parity hook: evaluates sampleEnvMap / pdfEnvMap (env_sampling.glsl:53-94) on the device for
 * n (u1,u2) pairs -> dir_pdf[4n] = (dir.xyz, pdf) and pdf_of_dir[n] = pdfEnvMap(dir). */
int ohb_env_sample_batch(ohb_ctx*, const float* u12, uint32_t n, float* dir_pdf, float* pdf_of_dir);
/* Parity hook: pdfEnvMap (env_sampling.glsl:79-94) for n caller-supplied unit directions dirs3[3n]. */
int ohb_env_pdf_batch(ohb_ctx*, const float* dirs3, uint32_t n, float* pdf);

/* NRD front-end packing (shaders/includes/rt/nrd_frontend.glsl:11-41 nrdPackRadianceHitDist / nrdLinearToYCoCg /
 * nrdNormHitDist / nrdYCoCgToLinear, pt_raygen_offline.rgen:106-127 nrdPackNormalRoughness) evaluated on the device for n
 * items: rad_hd_vz_rough[6n] = (radiance.rgb, hitDist, viewZ, roughness), normal_rough[4n] = (unit normal, roughness) ->
 * packed_radiance[4n] = (Y, Co, Cg, normalised hit distance), packed_normal[4n] = R10G10B10A2 operands before quantisation,
 * unpacked_rgb[3n] = nrdYCoCgToLinear(packed_radiance.xyz).  AOV arithmetic only: NRD itself is a third-party library. */
int ohb_nrd_pack_batch(ohb_ctx*, const float* rad_hd_vz_rough, const float* normal_rough, uint32_t n,
                       float* packed_radiance, float* packed_normal, float* unpacked_rgb);

/* ---- hybrid-RT techniques: the other two consumers of the TLAS (rt_shadow_technique.cpp, rt_gi_technique.cpp) ---------- */
/* Push constants of shaders/rt/rt_shadow.rgen:13-19 (RTShadowTechnique::record, rt_shadow_technique.cpp:503-505). */
typedef struct ohb_hybrid_shadow_params {
    float light_dir[3]; float light_radius;    /* direction the light travels; angular radius (directional) or sphere radius (point) */
    float light_pos[3]; float light_range;
    uint32_t light_type;                       /* 0 directional, 1 point / spot */
    uint32_t sample_count;
    uint32_t _pad[2];
} ohb_hybrid_shadow_params;
/* Push constants of shaders/rt/rt_gi.rgen:18-23. */
typedef struct ohb_hybrid_gi_params { float light_pos[3]; float light_intensity; uint32_t sample_count; uint32_t frame_index; uint32_t _pad[2]; } ohb_hybrid_gi_params;
/* rt_shadow.rgen: soft-shadow mask (R8_UNORM, W*H bytes) of the context's resolution from the G-buffer world positions
 * (RGBA32F, all-zero texel = sky) and octahedron-encoded normals (2 floats per pixel in [0,1], encoding.glsl:15-22). */
int ohb_hybrid_shadow(ohb_ctx*, const float* gbuf_position_rgba, const float* gbuf_normal_rg, const ohb_hybrid_shadow_params*, uint8_t* shadow_mask_r8);
/* rt_gi.rgen + rt_gi.rchit: one-bounce diffuse GI with the temporal blend, RGBA16F out (W*H*4 fp16 bit patterns).
 * instance_materials_rgba = materials[gl_InstanceID] of rt_gi.rchit:10-12 (rgb albedo, a >= 0.5 static), one vec4 per instance. */
int ohb_hybrid_gi(ohb_ctx*, const float* gbuf_position_rgba, const float* gbuf_normal_rg, const float* gbuf_albedo_rgba, const float* gi_history_rgba,
                  const float* instance_materials_rgba, uint32_t ninstances, const ohb_hybrid_gi_params*, uint16_t* gi_out_rgba16f);

/* ---- acceleration structure: createBLAS + buildTLAS (rt_acceleration_structure.cpp:205-535) -- */
/* How ohb_set_instances' BLAS/TLAS interface is realised.  OHB_ACCEL_FLATTEN (default, the static fast path): every instance is
 * transformed into ONE world-space 8-wide BVH.  OHB_ACCEL_TWO_LEVEL: one object-space BLAS per instance (createBLAS) under a TLAS
 * over the instances' world boxes (buildTLAS); rays are mapped into object space on descent.  Changing the mode invalidates
 * the structure (call ohb_build_accel). */
enum { OHB_ACCEL_FLATTEN = 0, OHB_ACCEL_TWO_LEVEL = 1 };
int ohb_set_accel_mode(ohb_ctx*, int mode);
int ohb_build_accel(ohb_ctx*);                         /* LBVH (Morton + radix sort) + SAH treelets */
/* VK_BUILD_ACCELERATION_STRUCTURE_MODE_UPDATE_KHR (rt_acceleration_structure.cpp:467 ALLOW_UPDATE, :503-512): the same n instances
 * with new object->world transforms.  Two-level: TLAS boxes refit, BLASes untouched.  Flattened: triangles re-transformed and the
 * tree refit bottom-up with its topology kept (no sort, no hierarchy build, no treelet pass).  Anything else changed -> rebuild. */
int ohb_update_instances(ohb_ctx*, const ohb_instance* instances, uint32_t n);
int ohb_get_accel_stats(ohb_ctx*, ohb_accel_stats*);

/* ---- per-frame: IRTRendererProfile::{setRenderSettings,setRenderSeed,resetAccumulation,
 *      notifyViewChanged,render} (rt_profile_renderer.hpp:150-175, path_tracer_render.cpp:38) ---- */
int      ohb_set_settings(ohb_ctx*, const ohb_settings*);
int      ohb_get_settings(ohb_ctx*, ohb_settings*);
void     ohb_set_seed(ohb_ctx*, uint32_t seed);        /* setRenderSeed: also resets accumulation */
void     ohb_reset_accumulation(ohb_ctx*);             /* m_sampleIndex = seed, history = 0       */
void     ohb_notify_view_changed(ohb_ctx*);
uint32_t ohb_frame_index(const ohb_ctx*);              /* getFrameIndex(): current sample index   */
/* One call == `nsamples` consecutive PathTracer::render() calls with this view/proj
 * (offline: sample indices advance by nsamples, one spp each; realtime: nsamples FRAMES of
 * settings.samples_per_frame spp each — ReSTIR GI temporal + spatial reuse, reprojected EMA, a-trous).
 * Asynchronous w.r.t. the host like vkQueueSubmit; readbacks synchronise.  Offline batches may run as two lanes on two
 * internal streams (OHB_LANES); the lanes join on the context's stream before the call returns, and the accumulation image is
 * bit-identical to the one-lane result (samples are folded in index order). */
int ohb_render(ohb_ctx*, const float view[16], const float proj[16], uint32_t nsamples);
/* Restrict offline rendering to a pixel rectangle (multi-GPU tile sharding, §8e). Default: full frame. */
int ohb_set_tile(ohb_ctx*, uint32_t x0, uint32_t y0, uint32_t w, uint32_t h);

/* ---- readback: getPixels / readbackHDRBuffers (renderer.cpp:700-760, render_dispatch.cpp:536-629) */
int ohb_read_ldr(ohb_ctx*, uint8_t* rgba8);                              /* W*H*4, top row first */
int ohb_read_hdr(ohb_ctx*, float* accum_rgba, float* albedo_rgba, float* normal_rgba); /* any may be NULL */
int ohb_synchronize(ohb_ctx*);
/* Device pointer + byte size of the RGBA32F accumulation image, for an in-place NCCL reduce of
 * sharded renders (§8e).  In sum mode (ohb_set_accum_mode(ctx,1)) the image holds
 * (sum.rgb, count) instead of the running mean so partial images add exactly. */
void*  ohb_accum_dev_ptr(ohb_ctx*, size_t* bytes);
int    ohb_set_accum_mode(ohb_ctx*, int sum_mode);
int    ohb_resolve(ohb_ctx*);   /* sum mode: divide by count + tonemap into the LDR image */
/* Zero the accumulation image.  A sum-mode render only overwrites the pixels of the tiles it renders
 * (history 0) and adds to them afterwards; a context reused for a second sharded image must start from
 * zero, or pixels outside this rank's tiles would carry the previous image into the next reduce.  (The
 * Vulkan path has no counterpart: PathTracer::resetAccumulation relies on frame 0 overwriting the
 * whole image, path_tracer.cpp:189-211.) */
int    ohb_clear_accum(ohb_ctx*);

/* ---- parity / measurement hooks ---------------------------------------------------------- */
/* Closest-hit and any-hit queries on caller-supplied rays (host buffers, copies included). */
int ohb_trace_batch(ohb_ctx*, const ohb_ray* rays, uint32_t n, ohb_hit* hits);
int ohb_occluded_batch(ohb_ctx*, const ohb_ray* rays, uint32_t n, uint8_t* occluded);
/* Per-sample radiance of the NEXT ohb_render call is also written here (n = W*H*nsamples*4
 * floats, sample-major), so tests can diff single samples against the oracle. NULL disables. */
int ohb_set_sample_dump(ohb_ctx*, float* host_rgba, size_t capacity_floats);
int ohb_get_counters(ohb_ctx*, ohb_counters*);
void ohb_reset_counters(ohb_ctx*);
/* Device time (ms, CUDA events on the context's stream) spent in traversal kernels / all kernels
 * since the last ohb_reset_counters, for the roofline line of bench.py. */
int ohb_get_timing(ohb_ctx*, float* trace_ms, float* shade_ms, float* total_ms);
int ohb_enable_timing(ohb_ctx*, int enable);
/* Per-category device time and launch count: [0] closest-hit traversal, [1] bounce (raygen body: NEE, MIS,
 * lobe sampling), [2] any-hit traversal, [3] film, [4] surface (closest-hit / miss shaders), [5] realtime per-pixel pass
 * (ReSTIR GI + EMA), [6] SVGF denoiser, [7] hit/miss queue sort (k_sort_hits). */
int ohb_get_timing_detail(ohb_ctx*, float ms[8], uint64_t launches[8]);
/* Realtime-profile parity hooks.  ohb_set_realtime_dump: the next ohb_render also copies, per pixel, the N-spp mean
 * radiance after the x0.75 clamp (pt_raygen_realtime.rgen:1552-1556), the diffuse ReSTIR GI term (:1755-1764) and the
 * a-trous output (:1850-1912) into the given host buffers (W*H*4 floats each, any may be NULL).
 * ohb_read_realtime_state: the reservoir planes written by the last frame (bindings 32-34: (x_s, M) (n_s, W) (Lo, valid))
 * and the surface / shading history (bindings 14/16). */
int ohb_set_realtime_dump(ohb_ctx*, float* radiance_rgba, float* gi_rgba, float* denoised_rgba);
int ohb_read_realtime_state(ohb_ctx*, float* res0, float* res1, float* res2, float* surface_history, float* shading_history);
/* AtrousDenoiser::dispatch (atrous_denoise.hpp:19-61, atrous_denoise.cpp:392-537) on caller-supplied images of the
 * context's resolution (realtime-profile context): RGBA8 beauty in / denoised out, RGBA32F normal AOV (N*0.5+0.5),
 * R32F linear view Z (1e30 = background), RG16F motion vectors packed x | y << 16 (pixel units, current - previous).
 * The denoiser's history lives in the context and ping-pongs like AtrousDenoiser's; reset != 0 discards it.
 * ohb_render runs the same kernels on its own images when settings.denoise_mode == OHB_DENOISE_ATROUS. */
int ohb_svgf_dispatch(ohb_ctx*, uint8_t* beauty_rgba8, const float* normal_rgba, const float* depth, const uint32_t* motion_rg16f, int reset);
/* SVGF denoiser parity hook (DenoiseMode::Atrous, AtrousDenoiser history images of atrous_denoise.cpp:79-121): the
 * history written by the last frame as raw fp16 bits — colour RGBA16F, moments (m1, m2, len, 0) RGBA16F, geometry
 * (viewZ, n) RGBA16F, W*H*4 uint16 each — and the guide AOVs of that frame: motion vectors RG16F packed x | y << 16
 * (binding 19) and linear view Z R32F (binding 20).  Any pointer may be NULL. */
int ohb_read_denoise_state(ohb_ctx*, uint16_t* color, uint16_t* moments, uint16_t* geom, uint32_t* motion, float* depth);
/* Whole-region device timer: two CUDA events recorded on the context's stream (the stream every
 * kernel of this context is launched on).  ohb_timer_stop synchronises and returns the elapsed ms. */
int ohb_timer_start(ohb_ctx*);
int ohb_timer_stop(ohb_ctx*, float* elapsed_ms);

#ifdef __cplusplus
}
#endif
#endif /* OHAO_B200_H */
