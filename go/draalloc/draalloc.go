// Package draalloc binds libdra_alloc.so (include/dra_alloc.h) with cgo.
//
// UNBUILT: the image this repository is developed in has no Go toolchain (`go version`: command not found),
// so this file has never been compiled.  It is the binding a maintainer of NVIDIA/k8s-dra-driver would add
// next to cmd/nvidia-dra-controller to put the B200 allocation path behind the classic
// controller.Driver surface (Allocate / UnsuitableNodes / Deallocate).  The C++ layer in
// k8s-dra-driver_b200/csrc/dra_host.cpp is the same logic, built and tested (tests/cpp/driver_test.cpp).
package draalloc

/*
#cgo CFLAGS: -I${SRCDIR}/../../include
#cgo LDFLAGS: -L${SRCDIR}/../../k8s-dra-driver_b200 -ldra_alloc
#include <stdlib.h>
#include "dra_alloc.h"
*/
import "C"

import (
	"errors"
	"fmt"
	"runtime"
	"sync"
	"unsafe"
)

// Records: identical layout to the C structs (spec/ALLOCATION.md §1).
type GpuRec struct {
	Busy       uint16
	Flags      uint8
	Model      uint8
	MemFreeMiB uint32
	Node       uint32
	ShareCnt   uint16
	_          uint16
}

type ClaimRec struct {
	Kind        uint8
	Profile     uint8
	Count       uint16
	Node        uint32
	MemLimitMiB uint32
	Group       uint32
}

type OutRec struct {
	Gpu     uint32
	Start   uint8
	Size    uint8
	Profile uint8
	Status  uint8
}

const (
	KindGpu    = 0
	KindMig    = 1
	KindShared = 2

	GpuMigEnabled    = 0x01
	GpuFullAllocated = 0x02
	GpuUnavailable   = 0x04

	StOK          = 0
	StPod         = 6 // pod mode: the pod does not fit on the node
	StSearchLimit = 7

	FlagExhaustive = 0x4 // DRA_F_EXHAUSTIVE: backtracking placement search (spec §12)
	CfgResident    = 0x8 // DRA_CFG_RESIDENT: resident kernel + doorbell for small batches
)

// Context owns one dra_ctx (one CUDA device, one stream, one inventory).  Calls are serialised, as the
// reference serialises its own (cmd/nvidia-dra-plugin/driver.go:119-120).
type Context struct {
	mu sync.Mutex
	h  *C.dra_ctx
}

func lastError(h *C.dra_ctx) error { return errors.New(C.GoString(C.dra_last_error(h))) }

func NewContext(device int, cfgFlags uint32) (*Context, error) {
	cfg := C.dra_cfg{abi_version: C.DRA_ABI_VERSION, device: C.int32_t(device), flags: C.uint32_t(cfgFlags)}
	var h *C.dra_ctx
	if rc := C.dra_ctx_create(&cfg, &h); rc != 0 {
		return nil, fmt.Errorf("dra_ctx_create: %w", lastError(nil))
	}
	c := &Context{h: h}
	runtime.SetFinalizer(c, func(c *Context) { c.Close() })
	return c, nil
}

func (c *Context) Close() {
	c.mu.Lock()
	defer c.mu.Unlock()
	if c.h != nil {
		C.dra_ctx_destroy(c.h)
		c.h = nil
	}
}

// SetPlacementTable mirrors what deviceLib.getGpuInfo collects per GPU model (nvlib.go:244-295).
func (c *Context) SetPlacementTable(model uint32, sizes [16]uint8, startMasks [16]uint16) error {
	var t C.dra_profile_tbl
	for p := 0; p < 16; p++ {
		t.ent[p].size = C.uint8_t(sizes[p])
		t.ent[p].start_mask = C.uint16_t(startMasks[p])
	}
	c.mu.Lock()
	defer c.mu.Unlock()
	if rc := C.dra_set_placement_table(c.h, C.uint32_t(model), &t); rc != 0 {
		return lastError(c.h)
	}
	return nil
}

func (c *Context) SetInventory(gpus []GpuRec, nodeOff []uint32) error {
	c.mu.Lock()
	defer c.mu.Unlock()
	var gp *C.dra_gpu_rec
	if len(gpus) > 0 {
		gp = (*C.dra_gpu_rec)(unsafe.Pointer(&gpus[0]))
	}
	rc := C.dra_set_inventory(c.h, gp, C.uint32_t(len(gpus)),
		(*C.uint32_t)(unsafe.Pointer(&nodeOff[0])), C.uint32_t(len(nodeOff)-1))
	if rc != 0 {
		return lastError(c.h)
	}
	return nil
}

// AllocateBatch = controller.Driver.Allocate batched over pods: every claim carries its selectedNode.
// outOff may be nil when every claim wants exactly one device.  cgo blocks the calling OS thread for the
// duration of the call (tens of microseconds per 10k claims), so keep calls batch-sized.
func (c *Context) AllocateBatch(claims []ClaimRec, outOff []uint32, nOut int) ([]OutRec, error) {
	out := make([]OutRec, nOut)
	if len(claims) == 0 {
		return out, nil
	}
	var oo *C.uint32_t
	if outOff != nil {
		oo = (*C.uint32_t)(unsafe.Pointer(&outOff[0]))
	}
	c.mu.Lock()
	defer c.mu.Unlock()
	rc := C.dra_allocate_batch(c.h, (*C.dra_claim_rec)(unsafe.Pointer(&claims[0])), C.uint32_t(len(claims)),
		oo, (*C.dra_out_rec)(unsafe.Pointer(&out[0])), C.uint32_t(nOut), 0)
	if rc != 0 {
		return nil, lastError(c.h)
	}
	return out, nil
}

// UnsuitableBatch = controller.Driver.UnsuitableNodes batched over pods; bit k of the result is set when
// the k-th (pod, candidate node) pair is suitable.
func (c *Context) UnsuitableBatch(claims []ClaimRec, podOff, candNodes, candOff []uint32, flags uint32) ([]byte, error) {
	nPod := len(podOff) - 1
	bits := make([]byte, (int(candOff[nPod])+7)/8+1)
	c.mu.Lock()
	defer c.mu.Unlock()
	var cp *C.dra_claim_rec
	if len(claims) > 0 {
		cp = (*C.dra_claim_rec)(unsafe.Pointer(&claims[0]))
	}
	var cn *C.uint32_t
	if len(candNodes) > 0 {
		cn = (*C.uint32_t)(unsafe.Pointer(&candNodes[0]))
	}
	rc := C.dra_unsuitable_batch(c.h, cp, C.uint32_t(len(claims)),
		(*C.uint32_t)(unsafe.Pointer(&podOff[0])), C.uint32_t(nPod), cn,
		(*C.uint32_t)(unsafe.Pointer(&candOff[0])), (*C.uint8_t)(unsafe.Pointer(&bits[0])), C.uint32_t(flags))
	if rc != 0 {
		return nil, lastError(c.h)
	}
	return bits, nil
}

// AllocatePodsBatch = Allocate with pod boundaries (spec §12): every pod (claims podOff[p]..podOff[p+1], one node)
// is placed atomically; with FlagExhaustive its MIG claims are placed by the backtracking search over every valid
// placement — what the classic mig.allocate did after UnsuitableNodes had found an assignment.
func (c *Context) AllocatePodsBatch(claims []ClaimRec, podOff, outOff []uint32, nOut int, flags uint32) ([]OutRec, error) {
	out := make([]OutRec, nOut)
	if len(claims) == 0 {
		return out, nil
	}
	var oo *C.uint32_t
	if outOff != nil {
		oo = (*C.uint32_t)(unsafe.Pointer(&outOff[0]))
	}
	c.mu.Lock()
	defer c.mu.Unlock()
	rc := C.dra_allocate_pods_batch(c.h, (*C.dra_claim_rec)(unsafe.Pointer(&claims[0])), C.uint32_t(len(claims)),
		(*C.uint32_t)(unsafe.Pointer(&podOff[0])), C.uint32_t(len(podOff)-1),
		oo, (*C.dra_out_rec)(unsafe.Pointer(&out[0])), C.uint32_t(nOut), C.uint32_t(flags))
	if rc != 0 {
		return nil, lastError(c.h)
	}
	return out, nil
}

func (c *Context) DeallocateBatch(claims []ClaimRec, outOff []uint32, out []OutRec) error {
	if len(claims) == 0 {
		return nil
	}
	var oo *C.uint32_t
	if outOff != nil {
		oo = (*C.uint32_t)(unsafe.Pointer(&outOff[0]))
	}
	c.mu.Lock()
	defer c.mu.Unlock()
	rc := C.dra_deallocate_batch(c.h, (*C.dra_claim_rec)(unsafe.Pointer(&claims[0])), C.uint32_t(len(claims)),
		oo, (*C.dra_out_rec)(unsafe.Pointer(&out[0])), C.uint32_t(len(out)))
	if rc != 0 {
		return lastError(c.h)
	}
	return nil
}

// ---- one global batch over the GPUs of a box (one Context per GPU of ONE process) --------------------------------------
// Set-up once, like the inventory: every context loads the WHOLE inventory, then
//   ctx[r].CommInitLocal(r, world); ctx[r].SetShardMap(bounds, strayRank); ctx[r].ShardExport(nOutMax);
//   ctx[r].PeerImportLocal(ctxs)
// Per batch: AllocateBatchGlobal on every context (same claims, without waiting for each other), then GatherRead on the one
// whose table is wanted.  The host does no partitioning and no merging (INTEGRATION.md §4).

func (c *Context) CommInitLocal(rank, world int) error {
	if rc := C.dra_comm_init_local(c.h, C.int(rank), C.int(world)); rc != 0 {
		return lastError(c.h)
	}
	return nil
}

// SetShardMap: rank r serves nodes [bounds[r], bounds[r+1]); len(bounds) == world+1.
func (c *Context) SetShardMap(bounds []uint32, strayRank int) error {
	if rc := C.dra_set_shard_map(c.h, (*C.uint32_t)(unsafe.Pointer(&bounds[0])), C.int(strayRank)); rc != 0 {
		return lastError(c.h)
	}
	return nil
}

func (c *Context) ShardExport(nOutMax uint32) error {
	if rc := C.dra_shard_export(c.h, C.uint32_t(nOutMax), 0, nil); rc != 0 {
		return lastError(c.h)
	}
	return nil
}

func (c *Context) PeerImportLocal(ctxs []*Context) error {
	hs := make([]*C.dra_ctx, len(ctxs))
	for i, o := range ctxs {
		hs[i] = o.h
	}
	if rc := C.dra_peer_import_local(c.h, (**C.dra_ctx)(unsafe.Pointer(&hs[0]))); rc != 0 {
		return lastError(c.h)
	}
	return nil
}

// AllocateBatchGlobal: claims must live in memory from dra_host_alloc (pinned): the device-side range filter then reads them
// over PCIe itself.  Returns at once; GatherRead synchronises.
func (c *Context) AllocateBatchGlobal(pinnedClaims unsafe.Pointer, nClaim int, nOut int, flags uint32) error {
	c.mu.Lock()
	defer c.mu.Unlock()
	rc := C.dra_allocate_batch_global_device(c.h, (*C.dra_claim_rec)(pinnedClaims), C.uint32_t(nClaim), nil, C.uint32_t(nOut), C.uint32_t(flags))
	if rc != 0 {
		return lastError(c.h)
	}
	return nil
}

func (c *Context) GatherRead(out []OutRec) error {
	c.mu.Lock()
	defer c.mu.Unlock()
	if rc := C.dra_gather_read(c.h, (*C.dra_out_rec)(unsafe.Pointer(&out[0])), C.uint32_t(len(out))); rc != 0 {
		return lastError(c.h)
	}
	return nil
}

// PeerRendezvous lines the contexts' streams up on the device (optional; dra_peer_rendezvous_device).
func (c *Context) PeerRendezvous() error {
	if rc := C.dra_peer_rendezvous_device(c.h); rc != 0 {
		return lastError(c.h)
	}
	return nil
}
