package draalloc

// UNBUILT (no Go toolchain in the development image).  Shape of the classic-DRA controller.Driver methods
// over the cgo binding; the tested equivalent is dra::Driver in k8s-dra-driver_b200/csrc/dra_host.cpp.

import "fmt"

// ClaimAllocation mirrors controller.ClaimAllocation as far as this path needs it.
type ClaimAllocation struct {
	ClaimUID        string
	IsMig           bool
	Count           uint32 // GpuClaimParameters.spec.count
	Profile         string // MigDeviceClaimParameters.spec.profile, e.g. "1g.5gb"
	GpuClaimName    string // co-location key
	MemLimitMiB     uint32 // MPS pinned-memory limit, limit.Megabyte arithmetic (sharing.go:234-237)
	Shared          bool
	UnsuitableNodes []string
	Devices         []string // canonical names, deviceinfo.go:74-80
	Error           error
}

type Driver struct {
	ctx       *Context
	nodeIndex map[string]uint32
	nodeOff   []uint32
	localIdx  []uint32          // GpuInfo.index per global GPU
	profEnum  map[string]uint8  // profile name -> NVML GI enum
	profID    map[uint8]int     // GI enum -> NVML profile id (device names)
	nextGroup uint32
	// Exhaustive: evaluate every pod atomically, MIG claims by the backtracking placement search (spec §12)
	Exhaustive bool
}

// groupAdjacent reorders a pod's claims so that members of a co-location group (same GpuClaimName) are adjacent,
// in order of the group's first appearance: the device only co-locates CONSECUTIVE claims of a group (spec §6).
func groupAdjacent(claims []*ClaimAllocation) []*ClaimAllocation {
	var runs [][]*ClaimAllocation
	runOf := map[string]int{}
	for _, ca := range claims {
		if ca.IsMig && ca.GpuClaimName != "" {
			if r, ok := runOf[ca.GpuClaimName]; ok {
				runs[r] = append(runs[r], ca)
				continue
			}
			runOf[ca.GpuClaimName] = len(runs)
		}
		runs = append(runs, []*ClaimAllocation{ca})
	}
	var out []*ClaimAllocation
	for _, r := range runs {
		out = append(out, r...)
	}
	return out
}

func (d *Driver) lower(ca *ClaimAllocation, node uint32, groups map[string]uint32) (ClaimRec, bool) {
	r := ClaimRec{Node: node, Count: 1}
	switch {
	case ca.IsMig:
		p, ok := d.profEnum[ca.Profile]
		if !ok {
			ca.Error = fmt.Errorf("unknown MIG profile %q", ca.Profile)
			return r, false
		}
		r.Kind, r.Profile = KindMig, p
		if ca.GpuClaimName != "" {
			g, ok := groups[ca.GpuClaimName]
			if !ok {
				d.nextGroup++
				g = d.nextGroup
				groups[ca.GpuClaimName] = g
			}
			r.Group = g
		}
	case ca.Shared:
		r.Kind, r.MemLimitMiB = KindShared, ca.MemLimitMiB
	default:
		if ca.Count < 1 || ca.Count > 32 { // DRA_MAX_COUNT; the device would answer INVALID with ONE slot
			ca.Error = fmt.Errorf("count must be in [1, 32]")
			return r, false
		}
		r.Kind, r.Count = KindGpu, uint16(ca.Count)
	}
	return r, true
}

// Allocate — controller.Driver.Allocate(ctx, claims, selectedNode).
func (d *Driver) Allocate(claims []*ClaimAllocation, selectedNode string) {
	node, ok := d.nodeIndex[selectedNode]
	groups := map[string]uint32{}
	var recs []ClaimRec
	var off []uint32
	var owner []*ClaimAllocation
	nOut := 0
	for _, ca := range groupAdjacent(claims) {
		if !ok {
			ca.Error = fmt.Errorf("unknown node %q", selectedNode)
			continue
		}
		if r, good := d.lower(ca, node, groups); good {
			recs, off, owner = append(recs, r), append(off, uint32(nOut)), append(owner, ca)
			if r.Kind == KindGpu {
				nOut += int(r.Count)
			} else {
				nOut++
			}
		}
	}
	var out []OutRec
	var err error
	if d.Exhaustive { // the pod = the claims of this call
		out, err = d.ctx.AllocatePodsBatch(recs, []uint32{0, uint32(len(recs))}, off, nOut, FlagExhaustive)
	} else {
		out, err = d.ctx.AllocateBatch(recs, off, nOut)
	}
	for i, ca := range owner {
		if err != nil {
			ca.Error = err
			continue
		}
		end := nOut
		if i+1 < len(off) {
			end = int(off[i+1])
		}
		for _, o := range out[off[i]:end] {
			if o.Status != StOK {
				ca.Error = fmt.Errorf("allocation failed: status %d", o.Status)
				ca.Devices = nil
				break
			}
			idx := d.localIdx[o.Gpu]
			if o.Profile < 16 {
				ca.Devices = append(ca.Devices, fmt.Sprintf("gpu-%d-mig-%d-%d-%d", idx, d.profID[o.Profile], o.Start, o.Size))
			} else {
				ca.Devices = append(ca.Devices, fmt.Sprintf("gpu-%d", idx))
			}
		}
	}
}

// UnsuitableNodes — controller.Driver.UnsuitableNodes(ctx, pod, claims, potentialNodes).
func (d *Driver) UnsuitableNodes(claims []*ClaimAllocation, potentialNodes []string) error {
	groups := map[string]uint32{}
	var recs []ClaimRec
	for _, ca := range groupAdjacent(claims) {
		r, ok := d.lower(ca, 0, groups)
		if !ok {
			for _, ca2 := range claims {
				ca2.UnsuitableNodes = append(ca2.UnsuitableNodes, potentialNodes...)
			}
			return nil
		}
		recs = append(recs, r)
	}
	cand := make([]uint32, len(potentialNodes))
	for i, n := range potentialNodes {
		if idx, ok := d.nodeIndex[n]; ok {
			cand[i] = idx
		} else {
			cand[i] = 0xFFFFFFFF
		}
	}
	var flags uint32
	if d.Exhaustive {
		flags = FlagExhaustive
	}
	bits, err := d.ctx.UnsuitableBatch(recs, []uint32{0, uint32(len(recs))}, cand, []uint32{0, uint32(len(cand))}, flags)
	if err != nil {
		return err
	}
	for k, n := range potentialNodes {
		if bits[k>>3]>>(uint(k)&7)&1 == 0 {
			for _, ca := range claims {
				ca.UnsuitableNodes = append(ca.UnsuitableNodes, n)
			}
		}
	}
	return nil
}
