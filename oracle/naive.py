"""Second, independent restatement of spec/ALLOCATION.md in pure Python (small cases only).

TEST INFRASTRUCTURE ONLY.  Written from the spec text with different data structures than
oracle/dra_oracle.c (per-node dicts, slice *sets* instead of bit masks, input-order processing with
explicit per-node queues) so that the C oracle is cross-checked by something that shares none of its
code.  PARITY UNPINNED for the search.
"""
from __future__ import annotations

NONE = 0xFFFFFFFF


def _slots(c, n_node, have_off):
    if _invalid(c, n_node, have_off):
        return 1
    return c["count"] if c["kind"] == 0 else 1


def _invalid(c, n_node, have_off):
    if c["kind"] > 2 or c["node"] >= n_node:
        return True
    if c["kind"] == 0 and (c["count"] == 0 or c["count"] > 32 or (not have_off and c["count"] != 1)):
        return True
    if c["kind"] == 1 and c["profile"] >= 16:
        return True
    return False


def _oprof(c):
    return {0: 0xFF, 2: 0xFE}.get(c["kind"], c["profile"])


def allocate(gpus, node_off, table, claims, out_off=None):
    """gpus: list of dicts (busy, flags, model, mem_free_mib, share_cnt); table[m][p] = (size, start_mask);
    claims: list of dicts.  Returns (out list of tuples (gpu,start,size,profile,status), gpus_after)."""
    n_node = len(node_off) - 1
    have_off = out_off is not None
    G = []
    for g in gpus:
        G.append({"used": {i for i in range(16) if (g["busy"] >> i) & 1}, "mig": bool(g["flags"] & 1),
                  "full": bool(g["flags"] & 2), "unavail": bool(g["flags"] & 4), "model": g["model"],
                  "mem": g["mem_free_mib"], "share": g["share_cnt"]})
    sl = [_slots(c, n_node, have_off) for c in claims]
    base = list(out_off) if have_off else list(range(len(claims)))
    n_out = max([b + s for b, s in zip(base, sl)], default=0)
    out = [None] * n_out

    def fill(i, st):
        for k in range(sl[i]):
            out[base[i] + k] = (NONE, 0, 0, _oprof(claims[i]), st)

    def placements(g, prof):
        size, mask = table[g["model"]][prof]
        return size, [s for s in range(16) if (mask >> s) & 1]

    def fit(used, size, starts):
        for s in starts:
            if not (set(range(s, s + size)) & used):
                return s
        return None

    queues = {n: [] for n in range(n_node)}
    for i, c in enumerate(claims):
        if c["node"] >= n_node:
            fill(i, 5)
        else:
            queues[c["node"]].append(i)

    for n in range(n_node):
        g0, g1 = node_off[n], node_off[n + 1]
        q = queues[n]
        k = 0
        while k < len(q):
            i = q[k]
            c = claims[i]
            if _invalid(c, n_node, have_off):
                fill(i, 5)
                k += 1
                continue
            if c["kind"] == 1 and c["group"] != 0:
                e = k + 1
                while (e < len(q) and e - k < 32 and claims[q[e]]["kind"] == 1
                       and claims[q[e]]["group"] == c["group"]
                       and not _invalid(claims[q[e]], n_node, have_off)):
                    e += 1
                placed = False
                for gi in range(g0, g1):
                    g = G[gi]
                    if not g["mig"] or g["unavail"] or g["full"]:
                        continue
                    used = set(g["used"])
                    recs = []
                    for m in q[k:e]:
                        size, starts = placements(g, claims[m]["profile"])
                        s = fit(used, size, starts) if starts else None
                        if s is None:
                            recs = None
                            break
                        used |= set(range(s, s + size))
                        recs.append((gi, s, size, claims[m]["profile"], 0))
                    if recs is not None:
                        g["used"] = used
                        for m, r in zip(q[k:e], recs):
                            out[base[m]] = r
                        placed = True
                        break
                if not placed:
                    for m in q[k:e]:
                        fill(m, 3)
                k = e
                continue
            if c["kind"] == 0:
                el = [gi for gi in range(g0, g1)
                      if not (G[gi]["mig"] or G[gi]["full"] or G[gi]["unavail"]) and G[gi]["share"] == 0]
                if len(el) < c["count"]:
                    fill(i, 1)
                else:
                    for j, gi in enumerate(el[: c["count"]]):
                        G[gi]["full"] = True
                        out[base[i] + j] = (gi, 0, 0, 0xFF, 0)
            elif c["kind"] == 1:
                offers = False
                done = False
                for gi in range(g0, g1):
                    g = G[gi]
                    size, starts = placements(g, c["profile"])
                    if not g["mig"] or g["unavail"] or not starts:
                        continue
                    offers = True
                    if g["full"]:
                        continue
                    s = fit(g["used"], size, starts)
                    if s is not None:
                        g["used"] |= set(range(s, s + size))
                        out[base[i]] = (gi, s, size, c["profile"], 0)
                        done = True
                        break
                if not done:
                    fill(i, 1 if offers else 2)
            else:
                done = False
                for gi in range(g0, g1):
                    g = G[gi]
                    if g["mig"] or g["full"] or g["unavail"] or g["share"] == 0xFFFF:
                        continue
                    if g["mem"] < c["mem_limit_mib"]:
                        continue
                    g["mem"] -= c["mem_limit_mib"]
                    g["share"] += 1
                    out[base[i]] = (gi, 0, 0, 0xFE, 0)
                    done = True
                    break
                if not done:
                    fill(i, 4)
            k += 1
    after = []
    for g0_, g in zip(gpus, G):
        after.append({"busy": sum(1 << s for s in g["used"]),
                      "flags": (1 if g["mig"] else 0) | (2 if g["full"] else 0) | (4 if g["unavail"] else 0),
                      "model": g["model"], "mem_free_mib": g["mem"], "share_cnt": g["share"]})
    return out, after


# ---- spec §12: pod mode + exhaustive placement search ---------------------------------------------------------
EXH_BUDGET = 4096
MAX_POD = 32


class _Limit(Exception):
    pass


def _state(gpus):
    return [{"used": {i for i in range(16) if (g["busy"] >> i) & 1}, "mig": bool(g["flags"] & 1),
             "full": bool(g["flags"] & 2), "unavail": bool(g["flags"] & 4), "model": g["model"],
             "mem": g["mem_free_mib"], "share": g["share_cnt"]} for g in gpus]


def _dump(gpus, G):
    return [{"busy": sum(1 << s for s in g["used"]),
             "flags": (1 if g["mig"] else 0) | (2 if g["full"] else 0) | (4 if g["unavail"] else 0),
             "model": g["model"], "mem_free_mib": g["mem"], "share_cnt": g["share"]} for g in G]


def eval_pod(G, g0, g1, table, pod, n_node, exhaustive, have_off=True):
    """One pod on the GPUs G[g0:g1] (list of state dicts, mutated only on success).
    Returns (placed, per-claim list of record lists)."""
    import copy
    inv = [_invalid(c, n_node, have_off) for c in pod]
    sl = [_slots(c, n_node, have_off) for c in pod]

    def failed(st):
        return False, [[(NONE, 0, 0, _oprof(c), 5 if bad else st)] * n for c, bad, n in zip(pod, inv, sl)]

    if any(inv):
        return failed(6)
    W = copy.deepcopy(G[g0:g1])
    recs = [None] * len(pod)
    for i, c in enumerate(pod):                       # step 2: GPU / SHARED claims in order
        if c["kind"] == 0:
            el = [k for k, g in enumerate(W) if not (g["mig"] or g["full"] or g["unavail"]) and g["share"] == 0]
            if len(el) < c["count"]:
                return failed(6)
            for k in el[: c["count"]]:
                W[k]["full"] = True
            recs[i] = [(g0 + k, 0, 0, 0xFF, 0) for k in el[: c["count"]]]
        elif c["kind"] == 2:
            for k, g in enumerate(W):
                if g["mig"] or g["full"] or g["unavail"] or g["share"] == 0xFFFF or g["mem"] < c["mem_limit_mib"]:
                    continue
                g["mem"] -= c["mem_limit_mib"]
                g["share"] += 1
                recs[i] = [(g0 + k, 0, 0, 0xFE, 0)]
                break
            else:
                return failed(6)
    mig = [i for i, c in enumerate(pod) if c["kind"] == 1]
    budget = [0]
    choice = {}
    for i in mig:                                     # pre-check: no placement at all for this claim on the node as it stands
        c = pod[i]
        fits = False
        for g in W:
            size, mask = table[g["model"]][c["profile"]]
            if not g["mig"] or g["unavail"] or g["full"] or not mask:
                continue
            if any((mask >> s) & 1 and not (set(range(s, s + size)) & g["used"]) for s in range(16)):
                fits = True
                break
        if not fits:
            return failed(6)

    def options(level):
        c = pod[mig[level]]
        want = None
        if c["group"]:
            for j in range(level):
                if pod[mig[j]]["group"] == c["group"]:
                    want = choice[j][0]
                    break
        for k, g in enumerate(W):
            if want is not None and k != want:
                continue
            size, mask = table[g["model"]][c["profile"]]
            if not g["mig"] or g["unavail"] or g["full"] or not mask:
                continue
            for s in range(16):
                if (mask >> s) & 1 and not (set(range(s, s + size)) & g["used"]):
                    yield k, s, size

    def search(level):
        if level == len(mig):
            return True
        for k, s, size in options(level):
            if budget[0] == EXH_BUDGET:
                raise _Limit
            budget[0] += 1
            W[k]["used"] |= set(range(s, s + size))
            choice[level] = (k, s, size)
            if search(level + 1):
                return True
            W[k]["used"] -= set(range(s, s + size))
            if not exhaustive:
                return False
        return False

    try:
        ok = search(0)
    except _Limit:
        return failed(7)
    if not ok:
        return failed(6)
    for level, i in enumerate(mig):
        k, s, size = choice[level]
        recs[i] = [(g0 + k, s, size, pod[i]["profile"], 0)]
    G[g0:g1] = W
    return True, recs


def allocate_pods(gpus, node_off, table, claims, pod_off, exhaustive, out_off=None):
    """spec §12.  Returns (out list, gpus_after)."""
    n_node = len(node_off) - 1
    have_off = out_off is not None
    G = _state(gpus)
    sl = [_slots(c, n_node, have_off) for c in claims]
    base = list(out_off) if have_off else list(range(len(claims)))
    n_out = max([b + s for b, s in zip(base, sl)], default=0)
    out = [None] * n_out
    for p in range(len(pod_off) - 1):
        c0, c1 = pod_off[p], pod_off[p + 1]
        pod = claims[c0:c1]
        if not pod:
            continue
        node = pod[0]["node"]
        if len(pod) > MAX_POD or node >= n_node or any(c["node"] != node for c in pod):
            for i in range(c0, c1):
                for k in range(sl[i]):
                    out[base[i] + k] = (NONE, 0, 0, _oprof(claims[i]), 5)
            continue
        _, recs = eval_pod(G, node_off[node], node_off[node + 1], table, pod, n_node, exhaustive, have_off)
        for i, r in zip(range(c0, c1), recs):
            for k, rec in enumerate(r):
                out[base[i] + k] = rec
    return out, _dump(gpus, G)


def pod_fits_bruteforce(gpus, table, pod):
    """Third opinion for tiny pods of MIG claims without selectors: does ANY assignment exist?  Enumerates every
    (gpu, start) vector with itertools.product — no pruning, no order."""
    import itertools
    G = _state(gpus)
    opts = []
    for c in pod:
        o = []
        for k, g in enumerate(G):
            size, mask = table[g["model"]][c["profile"]]
            if not g["mig"] or g["unavail"] or g["full"]:
                continue
            o += [(k, s, size) for s in range(16) if (mask >> s) & 1]
        opts.append(o)
    for vec in itertools.product(*opts):
        used = [set(g["used"]) for g in G]
        ok = True
        grp = {}
        for c, (k, s, size) in zip(pod, vec):
            if c["group"] and grp.setdefault(c["group"], k) != k:
                ok = False
                break
            cells = set(range(s, s + size))
            if cells & used[k]:
                ok = False
                break
            used[k] |= cells
        if ok:
            return True
    return False
