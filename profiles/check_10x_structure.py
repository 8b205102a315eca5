"""Structure invariants of the 10x graph (36.6 M edges) built with the library's radix sort, checked on the device:
sortedness and stability of the three orders, permutation property, pointer arrays."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch as th
from dgl_kgat_b200.graph import DeviceGraph
from dgl_kgat_b200.synth import synth_ckg
t0 = time.time()
ck = synth_ckg("amazon-book-10x"); N, R, trip = ck.cfg.n_nodes, ck.cfg.n_relations, ck.triplets
dev = th.device("cuda", 0)
dg = DeviceGraph(N, R, trip, dev)
E = dg.n_edges
a = {k: v[:E].long() if not k.endswith("_ptr") else v.long() for k, v in dg.arrays.items()}
t = th.from_numpy(trip).to(dev).long()
eid = a["csr_eid"]
assert th.equal(th.sort(eid).values, th.arange(E, device=dev)), "csr_eid is not a permutation"
assert th.equal(t[eid, 0], a["csr_dst"]) and th.equal(t[eid, 1], a["csr_rel"]) and th.equal(t[eid, 2], a["csr_src"])
key = a["csr_dst"] * R + a["csr_rel"]
assert bool((key[1:] >= key[:-1]).all()), "CSR not sorted by (head, rel)"
same = key[1:] == key[:-1]
assert bool((eid[1:][same] > eid[:-1][same]).all()), "CSR sort not stable"
assert th.equal(a["csr_ptr"], th.searchsorted(a["csr_dst"], th.arange(N + 1, device=dev)))
pos = a["csc_pos"]
assert th.equal(th.sort(pos).values, th.arange(E, device=dev))
src = a["csr_src"][pos]
assert bool((src[1:] >= src[:-1]).all()) and bool((pos[1:][src[1:] == src[:-1]] > pos[:-1][src[1:] == src[:-1]]).all())
assert th.equal(a["csc_ptr"], th.searchsorted(src, th.arange(N + 1, device=dev)))
assert th.equal(a["csc_dst"], a["csr_dst"][pos]) and th.equal(a["csc_eid"], eid[pos])
bp = a["bk_pos"]
rel = a["csr_rel"][bp]
assert bool((rel[1:] >= rel[:-1]).all()) and bool((bp[1:][rel[1:] == rel[:-1]] > bp[:-1][rel[1:] == rel[:-1]]).all())
assert th.equal(a["rel_ptr"], th.searchsorted(rel, th.arange(R + 1, device=dev)))
assert th.equal(a["bk_head"], a["csr_dst"][bp]) and th.equal(a["bk_tail"], a["csr_src"][bp])
print("10x structure ok: N=%d E=%d, %.1f s" % (N, E, time.time() - t0))
