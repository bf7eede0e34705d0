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
