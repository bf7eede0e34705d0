/*
 * dra_alloc.h — C ABI of libdra_alloc.so, the B200 (sm_100a) allocation hot path.
 *
 * This is the drop-in boundary for ONE path of NVIDIA/k8s-dra-driver: the per-node GPU/MIG claim search
 * behind the classic-DRA controller surface  Allocate() / UnsuitableNodes() / Deallocate()  that
 * BASELINE.json's north_star names.  That surface is ABSENT from the reference snapshot
 * (cmd/nvidia-dra-controller/ holds only main.go, imex.go, types.go — SURVEY.md F1); each entry point
 * below cites the reference interface it stands in for or the reference code that fixes its semantics.
 * A Go driver binds these with cgo (INTEGRATION.md shows the stub); tests bind them with ctypes.
 *
 * Plain C: pointers + sizes, caller-owned buffers, int return codes, no callbacks, no torch types.
 * Semantics are normative in spec/ALLOCATION.md; oracle/ is the CPU checker, never linked here.
 *
 * Threading: one in-flight call per dra_ctx (the reference serialises the same way with a global mutex,
 * cmd/nvidia-dra-plugin/driver.go:119-120, device_state.go:129-130).  Distinct contexts are independent.
 */
#ifndef DRA_ALLOC_H
#define DRA_ALLOC_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DRA_ABI_VERSION      2u   /* 2: flags on dra_unsuitable_batch, pod mode, sharded global batch, resident mode */

#define DRA_MAX_GPUS_PER_NODE 32u   /* one warp lane per GPU of a node */
#define DRA_MAX_MODELS        16u
#define DRA_MAX_PROFILES      16u   /* NVML GPU_INSTANCE_PROFILE_COUNT is 10 (go-nvml const.go:764-765) */
#define DRA_MAX_COUNT         32u   /* devices per GPU claim; k8s caps results per claim at 32
                                       (vendor/k8s.io/api/resource/v1beta1/types.go:792) */
#define DRA_MAX_GROUP         32u   /* members of one co-location run (same cap) */
#define DRA_MAX_POD           32u   /* claims of one pod in pod mode (spec §12) */
#define DRA_EXH_BUDGET        4096u /* descents of the exhaustive placement search per pod evaluation (spec §12) */
#define DRA_GPU_NONE          0xFFFFFFFFu

/* ---- records (spec/ALLOCATION.md §1) ------------------------------------------------------------ */

/* One physical GPU.  Mirrors the fields of GpuInfo that matter to allocation
 * (cmd/nvidia-dra-plugin/deviceinfo.go:30-43: index, migEnabled, memoryBytes) plus the occupancy the
 * removed NodeAllocationState CRD carried.  busy bit i = memory slice i in use
 * (deviceinfo.go:199-204, go-nvml nvml.h:9761-9765). */
typedef struct dra_gpu_rec {
    uint16_t busy;
    uint8_t  flags;         /* DRA_GPU_* */
    uint8_t  model;         /* row of the placement table */
    uint32_t mem_free_mib;  /* shareable memory left (SHARED claims) */
    uint32_t node;          /* owning node index */
    uint16_t share_cnt;     /* SHARED claims currently on this GPU */
    uint16_t rsvd;
} dra_gpu_rec;

#define DRA_GPU_MIG_ENABLED    0x01u  /* nvlib.go:152 / :316-318 — full GPU xor MIG parent */
#define DRA_GPU_FULL_ALLOCATED 0x02u
#define DRA_GPU_UNAVAILABLE    0x04u

/* One request.  Flat form of the legacy GpuClaimParameters / MigDeviceClaimParameters
 * (shapes survive only in demo/specs/mig+mps/sharing-demo-parameters.yaml:22-41 and
 * demo/specs/selectors/parameters.yaml:7-27) and of today's DeviceRequest + DeviceClass + CEL
 * `profile == '...'` + matchAttribute parentUUID (demo/specs/quickstart/gpu-test4.yaml:19-44). */
typedef struct dra_claim_rec {
    uint8_t  kind;          /* DRA_KIND_* */
    uint8_t  profile;       /* NVML GI profile enum (kind MIG) */
    uint16_t count;         /* devices wanted (kind GPU), else ignored */
    uint32_t node;          /* selectedNode (Allocate) — overridden per candidate in UnsuitableNodes */
    uint32_t mem_limit_mib; /* kind SHARED: MPS pinned-memory limit, sharing.go:234-237 arithmetic */
    uint32_t group;         /* != 0: consecutive MIG claims with equal group share one parent GPU */
} dra_claim_rec;

#define DRA_KIND_GPU    0u
#define DRA_KIND_MIG    1u
#define DRA_KIND_SHARED 2u

/* One allocated device (or one failed slot).  Bijective with the reference's device names
 * gpu-<index> / gpu-<parent>-mig-<profileId>-<start>-<size> (deviceinfo.go:74-80) that
 * DeviceRequestAllocationResult.Device carries (k8s types.go:795-840). */
typedef struct dra_out_rec {
    uint32_t gpu;           /* global inventory index, DRA_GPU_NONE when not allocated */
    uint8_t  start;
    uint8_t  size;
    uint8_t  profile;       /* MIG: claim profile; 0xFF GPU; 0xFE SHARED */
    uint8_t  status;        /* DRA_ST_* */
} dra_out_rec;

#define DRA_PROFILE_GPU    0xFFu
#define DRA_PROFILE_SHARED 0xFEu

#define DRA_ST_OK          0u
#define DRA_ST_NO_CAPACITY 1u
#define DRA_ST_BAD_PROFILE 2u
#define DRA_ST_GROUP       3u
#define DRA_ST_MEM_LIMIT   4u
#define DRA_ST_INVALID     5u
#define DRA_ST_POD         6u     /* pod mode (spec §12): the pod does not fit on the node; nothing of it was taken */
#define DRA_ST_SEARCH_LIMIT 7u    /* pod mode: the exhaustive search ran out of budget (DRA_EXH_BUDGET descents) */

/* One (model, profile) cell of the placement table: what getGpuInfo() collects from
 * GetGpuInstanceProfileInfo + GetGpuInstancePossiblePlacements (nvlib.go:244-295). */
typedef struct dra_prof_ent {
    uint8_t  size;          /* memory slices */
    uint8_t  rsvd;
    uint16_t start_mask;    /* bit s: placement {start s, size} is possible; 0 = profile not offered */
} dra_prof_ent;

typedef struct dra_profile_tbl {
    dra_prof_ent ent[DRA_MAX_PROFILES];
} dra_profile_tbl;

/* Per-GPU attributes a selector can test — the attributes GpuInfo.GetDevice publishes
 * (cmd/nvidia-dra-plugin/deviceinfo.go:102-132), interned to integers by the host.  spec §10. */
typedef struct dra_gpu_attr {
    uint32_t mem_total_mib;
    uint32_t cc;            /* cudaComputeCapability: major << 8 | minor */
    uint32_t index;
    uint16_t product;       /* host-interned productName id */
    uint16_t driver_major;
} dra_gpu_attr;

/* One instruction of a selector program (postfix, boolean stack).  Flat form of the legacy
 * GpuClaimParameters.spec.selector trees (demo/specs/selectors/parameters.yaml:7-27) and of the CEL subset
 * in demo/specs/quickstart/gpu-test6.yaml:23-31. */
typedef struct dra_sel_ins {
    uint8_t  op;            /* DRA_SEL_* */
    uint8_t  attr;          /* DRA_ATTR_* */
    uint8_t  cmp;           /* DRA_CMP_* */
    uint8_t  rsvd;
    uint32_t value;
} dra_sel_ins;

#define DRA_SEL_MAX_INS 8u
typedef struct dra_selector { dra_sel_ins ins[DRA_SEL_MAX_INS]; } dra_selector;

#define DRA_SEL_END 0u
#define DRA_SEL_CMP 1u
#define DRA_SEL_AND 2u
#define DRA_SEL_OR  3u
#define DRA_SEL_NOT 4u
#define DRA_ATTR_MEMORY_MIB   0u
#define DRA_ATTR_CC           1u
#define DRA_ATTR_INDEX        2u
#define DRA_ATTR_PRODUCT      3u
#define DRA_ATTR_DRIVER_MAJOR 4u
#define DRA_CMP_EQ 0u
#define DRA_CMP_NE 1u
#define DRA_CMP_LT 2u
#define DRA_CMP_LE 3u
#define DRA_CMP_GT 4u
#define DRA_CMP_GE 5u
#define DRA_CMP_IN_MASK 6u

/* ---- context ------------------------------------------------------------------------------------- */

typedef struct dra_ctx dra_ctx;

typedef struct dra_cfg {
    uint32_t abi_version;   /* DRA_ABI_VERSION */
    int32_t  device;        /* CUDA device ordinal */
    void*    stream;        /* cudaStream_t to run on; NULL = the context creates its own */
    uint32_t max_claims;    /* capacity hint; buffers grow on demand */
    uint32_t flags;         /* DRA_CFG_* */
} dra_cfg;

#define DRA_CFG_USE_GRAPH  0x1u  /* dra_allocate_batch: replay H2D -> kernels -> D2H as ONE CUDA graph launch when a
                                    call repeats the previous call's buffers and sizes (captured on the 2nd such call) */
#define DRA_CFG_NO_FUSED   0x2u  /* never take the single-launch path (always bucket + pack); for tests */
#define DRA_CFG_NO_DIRECT  0x4u  /* dra_allocate_batch: never let the single-launch kernel read the claims from / write the
                                    results to the pinned host buffers itself (direct host I/O, one cooperative launch);
                                    always use copy-engine transfers around the kernels */

#define DRA_CFG_RESIDENT   0x8u  /* dra_allocate_batch: keep the single-launch kernel RESIDENT and hand it batches by doorbell (a command
                                    line + sequence number in mapped host memory, completion word back) — no kernel launch, no stream
                                    call per batch.  Applies to batches that fit the staged single-launch form on an inventory of at most
                                    (SMs - 1) nodes with page-locked buffers; everything else stops the resident kernel first and takes
                                    the usual path.  The kernel leaves by itself after DRA_SERVE_IDLE_MS (20) ms without a batch. */

/* error codes (negative) */
#define DRA_OK        0
#define DRA_E_INVAL  (-1)
#define DRA_E_CUDA   (-2)
#define DRA_E_NCCL   (-3)
#define DRA_E_NOMEM  (-4)
#define DRA_E_STATE  (-5)

/* batch flags */
#define DRA_F_NODE_SORTED 0x1u  /* claims already grouped by ascending node (stable): skip bucketing.
                                   Verified on the device; violation -> DRA_E_INVAL, nothing changed */
#define DRA_F_FRESH_INVENTORY 0x2u  /* evaluate against the inventory as dra_set_inventory last loaded it
                                   (ignoring earlier batches); the result becomes the live inventory */

#define DRA_F_EXHAUSTIVE 0x4u   /* pod mode (spec §12): backtracking search over every valid (GPU, placement) of the pod's MIG
                                   claims — a pod fails only if NO assignment exists.  dra_allocate_pods_batch and
                                   dra_unsuitable_batch; the enumeration it searches is the one deviceLib.getGpuInfo builds
                                   (cmd/nvidia-dra-plugin/nvlib.go:244-295) */

int  dra_abi_version(void);

/* Creates a context bound to one CUDA device + one stream.  Fails with DRA_E_CUDA when no usable
 * sm_100 device exists — there is no CPU fallback. */
int  dra_ctx_create(const dra_cfg* cfg, dra_ctx** out);
void dra_ctx_destroy(dra_ctx* ctx);
const char* dra_last_error(const dra_ctx* ctx);      /* ctx may be NULL: last create error */

/* Placement table for one GPU model — replaces the per-GPU migProfiles list built by
 * deviceLib.getGpuInfo (cmd/nvidia-dra-plugin/nvlib.go:244-295). */
int  dra_set_placement_table(dra_ctx* ctx, uint32_t model, const dra_profile_tbl* tbl);

/* Inventory of every node this context serves — replaces the AllocatableDevices map published per node
 * (cmd/nvidia-dra-plugin/nvlib.go:111-180, driver.go:71-83) / the per-node NodeAllocationState CRD of
 * the classic driver.  gpus sorted by (node, local index); node_off has n_node+1 entries. Copied. */
int  dra_set_inventory(dra_ctx* ctx, const dra_gpu_rec* gpus, uint32_t n_gpu,
                       const uint32_t* node_off, uint32_t n_node);
/* Optional, spec §10: attributes of the same n_gpu GPUs (same order) and the selector table.  A claim names
 * its selector by 1-based id: kinds GPU and MIG in mem_limit_mib, kind SHARED in group. */
int  dra_set_gpu_attrs(dra_ctx* ctx, const dra_gpu_attr* attrs, uint32_t n_gpu);
int  dra_set_selectors(dra_ctx* ctx, const dra_selector* sels, uint32_t n_sel);
/* Read the live inventory back (n_gpu records). */
int  dra_get_inventory(dra_ctx* ctx, dra_gpu_rec* gpus, uint32_t n_gpu);
/* Restore the live inventory to what dra_set_inventory last loaded (device-side copy). */
int  dra_reset_inventory(dra_ctx* ctx);

/* Allocate(): stands in for  controller.Driver.Allocate(ctx, claims []*ClaimAllocation, selectedNode)
 * (k8s.io/dynamic-resource-allocation/controller, not vendored: vendor/modules.txt:757-760) batched over
 * many pods — every claim carries its selectedNode.  Mutates the context's inventory.
 *   claims[n_claim]           host memory
 *   out_off[n_claim] or NULL  first OutRec slot of each claim (NULL: slot i, all counts must be 1)
 *   out[n_out]                host memory, filled in input order
 * Per-claim failure is NOT an error (out[].status != 0), like the per-claim Error strings of
 * cmd/nvidia-dra-plugin/driver.go:126-137. */
int  dra_allocate_batch(dra_ctx* ctx, const dra_claim_rec* claims, uint32_t n_claim,
                        const uint32_t* out_off, dra_out_rec* out, uint32_t n_out, uint32_t flags);

/* Same, all pointers are DEVICE memory on ctx's device; enqueued on ctx's stream, returns without
 * synchronising (errors detected on the device surface at the next dra_ctx_sync). */
int  dra_allocate_batch_device(dra_ctx* ctx, const dra_claim_rec* d_claims, uint32_t n_claim,
                               const uint32_t* d_out_off, dra_out_rec* d_out, uint32_t n_out,
                               uint32_t flags);
int  dra_ctx_sync(dra_ctx* ctx);

/* UnsuitableNodes(): stands in for  controller.Driver.UnsuitableNodes(ctx, pod, claims, potentialNodes)
 * batched over pods.  Pure: evaluates each pod's claims on a snapshot of each candidate node.
 *   pod_off[n_pod+1]   claim range of each pod
 *   cand_off[n_pod+1]  range of each pod's candidates in cand_nodes
 *   suitable_bits      ceil(cand_off[n_pod]/8) bytes; bit k = k-th (pod,candidate) pair is suitable
 * Dense form: cand_nodes == NULL and cand_off == NULL evaluates every pod against EVERY node of the inventory
 * (pair k = pod * n_node + node; suitable_bits has ceil(n_pod*n_node/8) bytes) without any candidate arrays. */
int  dra_unsuitable_batch(dra_ctx* ctx, const dra_claim_rec* claims, uint32_t n_claim,
                          const uint32_t* pod_off, uint32_t n_pod,
                          const uint32_t* cand_nodes, const uint32_t* cand_off,
                          uint8_t* suitable_bits, uint32_t flags /* 0 or DRA_F_EXHAUSTIVE (spec §12) */);

/* Allocate() with pod boundaries (spec §12): what the classic driver's Allocate did after UnsuitableNodes had found
 * an assignment for ALL of a pod's claims on the selected node — each pod (claims
 * pod_off[p] .. pod_off[p+1], all naming the same node, at most DRA_MAX_POD) is placed atomically; with
 * DRA_F_EXHAUSTIVE its MIG claims are placed by the backtracking search over the placement enumeration of
 * nvlib.go:244-295 (multi-request shape: demo/specs/quickstart/gpu-test4.yaml:19-44).  A pod that does not fit
 * takes nothing: its slots carry DRA_ST_POD (or DRA_ST_SEARCH_LIMIT).  Host buffers; flags: DRA_F_EXHAUSTIVE,
 * DRA_F_FRESH_INVENTORY. */
int  dra_allocate_pods_batch(dra_ctx* ctx, const dra_claim_rec* claims, uint32_t n_claim,
                             const uint32_t* pod_off, uint32_t n_pod,
                             const uint32_t* out_off, dra_out_rec* out, uint32_t n_out, uint32_t flags);

/* Deallocate(): stands in for controller.Driver.Deallocate(ctx, claim).  Inverse update from the
 * claims and the OutRecs dra_allocate_batch produced for them (same out_off convention). */
int  dra_deallocate_batch(dra_ctx* ctx, const dra_claim_rec* claims, uint32_t n_claim,
                          const uint32_t* out_off, const dra_out_rec* out, uint32_t n_out);

/* ---- multi-GPU (one process per GPU; nodes are sharded whole) ------------------------------------ */

/* 128-byte NCCL unique id, generated on rank 0 and handed to the other ranks by the host runtime. */
int  dra_comm_unique_id(void* id128);
int  dra_comm_init(dra_ctx* ctx, const void* id128, int rank, int world);
/* dra_allocate_batch_device on this rank's claims, then ONE ncclAllGather of n_per_rank OutRecs per rank
 * into d_out_all[world * n_per_rank] on the same stream (slots past this rank's n_out are zero-filled). */
int  dra_allocate_batch_gather_device(dra_ctx* ctx, const dra_claim_rec* d_claims, uint32_t n_claim,
                                      const uint32_t* d_out_off, dra_out_rec* d_out_all,
                                      uint32_t n_out, uint32_t n_per_rank, uint32_t flags);

/* The complete table of the last gather: device pointer (valid until the second-next gather call) and the
 * padded per-rank length; with the peer all-gather set up d_out_all may be NULL and this is the result. */
int  dra_gather_table(dra_ctx* ctx, const dra_out_rec** d_table, uint32_t* n_per_rank);
/* Device -> host read of the first n_rec records of that table on ctx's stream, then synchronises. */
int  dra_gather_read(dra_ctx* ctx, dra_out_rec* out_all, uint32_t n_rec);

/* Peer-memory all-gather (OutRecs as self-validating 16-byte packets over NVLink, no fence / flag; measurements in
 * profiles/peer_bench_r02.txt) — when set up, dra_allocate_batch_gather_device uses it instead of ncclAllGather.  Call after dra_comm_init on every rank:
 *   dra_peer_export(ctx, n_per_rank, handle)   allocates this rank's gather buffer, returns its 64-byte
 *                                              cudaIpcMemHandle_t;  the host runtime all-gathers the handles;
 *   dra_peer_import(ctx, handles)              world * 64 bytes, in rank order.
 * Requires all ranks on one node with P2P access; on failure the NCCL path stays in use.
 * dra_peer_export(ctx, 0, NULL) switches the peer path off again. */
int  dra_peer_export(dra_ctx* ctx, uint32_t n_per_rank, void* handle64);
int  dra_peer_import(dra_ctx* ctx, const void* handles);

/* The same mapping for contexts of ONE process (one Go driver process driving all GPUs of a box, or tests): instead of
 * NCCL + IPC handles, call dra_comm_init_local(ctx, rank, world) on every context, dra_peer_export / dra_shard_export
 * (handle may be NULL), then dra_peer_import_local(ctx, ctxs) with the world contexts in rank order (peer access between
 * their devices is enabled here).  Calls of different ranks must then be issued without waiting for each other.
 * One context per DEVICE is the supported layout; several ranks on one device work for batches that take the
 * single-launch kernel (the tests do it) but can starve each other on the sort path (a waiting gather holds the SMs). */
int  dra_comm_init_local(dra_ctx* ctx, int rank, int world);
int  dra_peer_import_local(dra_ctx* ctx, dra_ctx* const* ctxs);

/* Cross-rank rendezvous ON THE DEVICE: enqueues one tiny kernel on the context's stream that tells every peer "rank r is
 * here" (one 4-byte store over NVLink each) and waits until every peer has said the same (after dra_peer_import /
 * dra_peer_import_local; every rank must call it the same number of times; no host synchronisation, returns at once).
 * It orders nothing but time: the ranks' streams leave it within one NVLink latency of each other.  bench.py calls it
 * between its untimed L2 flush and the timed step, so that a rank's step does not absorb its peers' flush-time skew.
 * A peer that never arrives within DRA_PEER_TIMEOUT_MS sets the context's error flag (next call: DRA_E_STATE). */
int  dra_peer_rendezvous_device(dra_ctx* ctx);

/* ---- sharded GLOBAL batch (BASELINE configs[2]: one batch, claim x node space sharded over the GPUs of a box) -----
 * Every rank loads the WHOLE inventory (dra_set_inventory) and serves a contiguous node range
 * (dra_set_shard; pool = node, nodes are never split: vendor/k8s.io/dynamic-resource-allocation/kubeletplugin/
 * draplugin.go:427-435).  dra_allocate_batch_global_device takes the SAME global claim array on every rank: the rank
 * filters the claims of its nodes on the device (stable compaction in input order), allocates them, and writes each
 * OutRec at the claim's GLOBAL slot of every rank's result table — the path's one collective, an all-gather done as
 * 16-byte self-validating packets over NVLink inside the allocation kernel (no host partition / merge, no un-permute).
 * Claims that name no node are answered by the rank with take_stray != 0 (exactly one rank).
 *   dra_shard_export(ctx, n_out_max, cap_per_rank, handle)  allocates table + staging (cap_per_rank = upper bound of the
 *        OutRec slots one rank can own, 0 = n_out_max); handles are exchanged and mapped with dra_peer_import as above
 *   result: dra_gather_table / dra_gather_read (n_out records, input order, global GPU indices)
 * out_off must tile [0, n_out) (slots no claim owns keep their previous contents).  flags: DRA_F_FRESH_INVENTORY.
 * With world == 1 no export is needed (the table is local).  A rank's live inventory is authoritative for its own
 * node range only.  d_claims may also be a pinned host buffer (dra_host_alloc): the compaction kernel reads every claim
 * exactly once, so it then ingests the batch over PCIe itself — no copy-engine transfer in front of the call. */
int  dra_set_shard(dra_ctx* ctx, uint32_t node_lo, uint32_t node_hi, int take_stray);
/* The whole partition: rank r serves nodes [bounds[r], bounds[r+1]) (bounds has world+1 entries, bounds[world] = n_node),
 * stray_rank answers claims naming no node.  Since every rank reads the whole claim array, it then also COUNTS what every
 * rank will answer — receivers know how many records to expect from each peer before the first one arrives, and the gather
 * runs without any header or completion round.  Preferred over dra_set_shard whenever world > 1. */
int  dra_set_shard_map(dra_ctx* ctx, const uint32_t* bounds, int stray_rank);
int  dra_shard_export(dra_ctx* ctx, uint32_t n_out_max, uint32_t cap_per_rank, void* handle64);
int  dra_allocate_batch_global_device(dra_ctx* ctx, const dra_claim_rec* d_claims, uint32_t n_claim,
                                      const uint32_t* d_out_off, uint32_t n_out, uint32_t flags);

/* ---- adjacent integer searches of the reference, batched (API completeness) -------------- */

/* limit.Megabyte() over n quantities (api/nvidia.com/resource/gpu/v1alpha1/sharing.go:234-237):
 * mib[i] = bytes[i]/1024/1024 truncated toward zero, valid[i] = mib[i] > 0. */
int  dra_mps_limits_batch(dra_ctx* ctx, const int64_t* bytes, uint32_t n, int64_t* mib, uint8_t* valid);
/* imexDomainOffsets.add's search over n_dom domains (cmd/nvidia-dra-controller/imex.go:336-349): used offsets of
 * domain d are used[dom_off[d] .. dom_off[d+1]); out[d] = lowest multiple of step below limit not in them, -1 if none. */
int  dra_imex_offsets_batch(dra_ctx* ctx, const int32_t* used, const uint32_t* dom_off, uint32_t n_dom,
                            int32_t step, int32_t limit, int32_t* out);

/* One-time calibration (optional): measures, on this device and the loaded inventory, up to which batch size the
 * single-launch kernel beats the bucket + pack chain, and uses that crossover from then on (the built-in values were
 * measured on one B200 box).  The live inventory is preserved.  *crossover_claims (may be NULL) receives the result. */
int  dra_calibrate(dra_ctx* ctx, uint32_t* crossover_claims);

/* ---- resident mode (DRA_CFG_RESIDENT) ------------------------------------------------------------------------------
 * Optional explicit control: start the resident kernel now (the first eligible dra_allocate_batch would), stop it (every
 * other entry point of the context does so implicitly before it runs).  dra_serve_batches: batches it has answered. */
int  dra_serve_start(dra_ctx* ctx);
int  dra_serve_stop(dra_ctx* ctx);
uint64_t dra_serve_batches(const dra_ctx* ctx);

/* ---- host memory + instrumentation ---------------------------------------------------------------- */

/* Page-locked host buffers: passing these to the *_batch calls skips the internal staging copy. */
void* dra_host_alloc(size_t bytes);
void  dra_host_free(void* p);

/* Kernel launches issued by this context since creation (all of them this library's own kernels). */
uint64_t dra_launch_count(const dra_ctx* ctx);
/* When enabled, CUDA events bracket each kernel of the next calls; dra_get_timings returns the last
 * call's per-stage device times in microseconds: [0] bucket-hist [1] bucket-scan [2] bucket-scatter
 * [3] pack [4] all-gather.  Returns the number of floats written. */
int  dra_set_profiling(dra_ctx* ctx, int enabled);
/* Instrumentation (env DRA_TIMELINE=1): 8 clock stamps per CTA of the last single-launch kernel; returns the
 * number of u64 words copied to host. */
int  dra_debug_timeline(dra_ctx* ctx, unsigned long long* host, uint32_t n);
/* Instrumentation (env DRA_TIMELINE=1): globaltimer of the last shard compaction's first CTA in / last CTA out. */
int  dra_debug_shard_times(dra_ctx* ctx, unsigned long long* two);
/* Instrumentation: globaltimer stamps of the resident kernel's last batch — doorbell seen, CTA 0's egress issued, every CTA
 * fenced, completion word written.  Stops the resident kernel. */
int  dra_debug_serve_times(dra_ctx* ctx, unsigned long long* four);
/* Instrumentation: enqueue an empty kernel of the given shape on ctx's stream (launch-floor calibration). */
int  dra_debug_noop(dra_ctx* ctx, uint32_t grid, uint32_t block, uint32_t smem_bytes);
int  dra_get_timings(dra_ctx* ctx, float* us, int n);

#ifdef __cplusplus
}
#endif
#endif /* DRA_ALLOC_H */
