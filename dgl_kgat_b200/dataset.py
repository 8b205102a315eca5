"""Graph inputs with the reference's ``DataLoader`` surface (reference ``dataset.py``), DGL-free.

What is kept (consumers: kgat.py:74-80, :95-117, :147, :173, :185):
``n_users, n_items, n_train, n_valid, n_test, n_KG_entity, n_KG_relation,
n_train_KG_triplet, n_test_KG_triplet, train_KG_triplet, test_KG_triplet, item_id_range,
train_user_dict, valid_user_dict, test_user_dict, train_valid_user_dict, train_pairs,
valid_pairs, test_pairs, train_g, test_g, KG_sampler(), CF_pair_sampler()``.

Id layout (dataset.py:60-73): users ``[0, U)``, then items, then the other KG entities,
i.e. item and entity ids are offset by ``n_users``; the user->items dicts keep RAW item
ids (they are built before the offset, dataset.py:45-50) so their values index the
``(n_items,)`` score vector directly.  Triplets: ``[KG ; (u, R_kg, v) ; (v, R_kg+1, u)]``
(dataset.py:84-93, :165-185).  ``train_g`` / ``test_g`` are :class:`CKGraph` objects built
once and cached (the reference rebuilds a DGLGraph per access, dataset.py:112-130; the
observable content is the same).

On-disk format (dataset.py:14-18, :150-152; datasets/README.md): tab-separated files with
a header row, ``datasets/<name>/data/{uv_train,uv_val,uv_test}.pd`` (columns u, v) and
``kg_item.pd`` (columns h, r, t), ids consecutive from 0.  ``DataLoader.from_synthetic``
builds the same object from :mod:`synth` without touching the disk.
"""
from __future__ import annotations

import os
import random as rd

import numpy as np
import pandas as pd

from . import synth
from .graph import CKGraph

_UV_FILES = ["uv_train.pd", "uv_val.pd", "uv_test.pd"]
_KG_FILE = "kg_item.pd"
_SEP = "\t"


def _read_table(path):
    if not os.path.isfile(path):
        raise FileNotFoundError("%s does not exist" % path)
    return pd.read_csv(path, sep=_SEP, header=0, dtype=np.int32)


def _pairs_to_dict(u, v):
    """{user: array of item ids}, users ascending, items in stable input order
    (dataset.py:187-196 groups a frame that was stably sorted by u)."""
    u = np.asarray(u)
    v = np.asarray(v)
    order = np.argsort(u, kind="stable")
    us, vs = u[order], v[order]
    uniq, start = np.unique(us, return_index=True)
    bounds = np.append(start, len(us))
    return {uniq[i]: vs[bounds[i]: bounds[i + 1]] for i in range(len(uniq))}


class EdgeBatchSampler:
    """One shuffled pass over the edges of a (src -> dst) graph in batches, each positive
    edge paired with one uniformly corrupted copy.  Stands in for
    ``dgl.contrib.sampling.EdgeSampler(..., neg_sample_size=1, shuffle=True,
    return_false_neg=True)`` as used at dataset.py:223-233: ``negative_mode='head'``
    replaces the source node, ``'tail'`` the destination node, the replacement is uniform
    over all nodes, and ``false_neg`` flags corrupted edges that exist in the graph with
    the same relation."""

    def __init__(self, src, dst, etype, n_nodes, batch_size, negative_mode, rng):
        self.src, self.dst, self.n_nodes = np.asarray(src, np.int64), np.asarray(dst, np.int64), int(n_nodes)
        self.etype = None if etype is None else np.asarray(etype, np.int64)
        self.batch_size, self.mode, self.rng = int(batch_size), negative_mode, rng
        self._n_rel = 1 if self.etype is None else int(self.etype.max()) + 1
        self._keys = np.unique(self._edge_key(self.src, self.dst, self.etype))

    def _edge_key(self, s, d, r):
        k = s * self.n_nodes + d
        return k if r is None else k * self._n_rel + r

    def __iter__(self):
        E = self.src.shape[0]
        perm = self.rng.permutation(E)
        for b in range(0, E, self.batch_size):
            eid = perm[b: b + self.batch_size]
            s, d = self.src[eid], self.dst[eid]
            r = None if self.etype is None else self.etype[eid]
            corrupt = self.rng.integers(0, self.n_nodes, size=eid.shape[0])
            ns, nd = (corrupt, d) if self.mode == "head" else (s, corrupt)
            q = self._edge_key(ns, nd, r)                      # membership by binary search: _keys is sorted
            pos = np.minimum(np.searchsorted(self._keys, q), self._keys.shape[0] - 1)
            false_neg = (self._keys[pos] == q).astype(np.int64)
            yield s, d, r, ns, nd, false_neg


class DeviceEdgeBatchSampler:
    """:class:`EdgeBatchSampler` with the edge list, the shuffling, the corruption and the
    false-negative lookup resident on one torch device: batches are
    device tensors, so the training loop does no per-batch host->device copies (kgat.py:123-128,
    :153-158) and the host does no sampling work.  Same batch contract, torch's RNG."""

    def __init__(self, src, dst, etype, n_nodes, batch_size, negative_mode, device, generator=None):
        import torch as th
        self.th, self.device, self.gen = th, th.device(device), generator
        self.src = th.as_tensor(np.asarray(src), dtype=th.long, device=self.device)
        self.dst = th.as_tensor(np.asarray(dst), dtype=th.long, device=self.device)
        self.etype = None if etype is None else th.as_tensor(np.asarray(etype), dtype=th.long, device=self.device)
        self.n_nodes, self.batch_size, self.mode = int(n_nodes), int(batch_size), negative_mode
        self._n_rel = 1 if self.etype is None else int(self.etype.max().item()) + 1
        self._keys = th.unique(self._edge_key(self.src, self.dst, self.etype))      # sorted

    def _edge_key(self, s, d, r):
        k = s * self.n_nodes + d
        return k if r is None else k * self._n_rel + r

    # batches are drawn SUPER at a time (a handful of kernels over SUPER * batch_size edges) and handed
    # out as views, so an iteration of the training loop costs no sampling launches of its own
    SUPER = 64

    def __iter__(self):
        th = self.th
        E = self.src.shape[0]
        perm = th.randperm(E, device=self.device, generator=self.gen)
        span = self.batch_size * self.SUPER
        for c in range(0, E, span):
            eid = perm[c: c + span]
            s, d = self.src[eid], self.dst[eid]
            r = None if self.etype is None else self.etype[eid]
            corrupt = th.randint(0, self.n_nodes, (eid.shape[0],), device=self.device, generator=self.gen)
            ns, nd = (corrupt, d) if self.mode == "head" else (s, corrupt)
            q = self._edge_key(ns, nd, r)
            pos = th.searchsorted(self._keys, q).clamp_(max=self._keys.shape[0] - 1)
            false_neg = (self._keys[pos] == q).to(th.long)
            for b in range(0, eid.shape[0], self.batch_size):
                e = b + self.batch_size
                yield (s[b:e], d[b:e], None if r is None else r[b:e], ns[b:e], nd[b:e], false_neg[b:e])


class UniqueGroupSampler:
    """The reference's pure-Python fallback samplers (dataset.py:246-266 ``pos_mode="unique",
    neg_mode="tail"`` and :294-310 ``pos_mode="unique", neg_mode="exclude_pos"``) as array code on one
    torch device (CPU or CUDA): a batch holds ``batch_size`` DISTINCT group keys (heads / users; drawn with
    replacement only when there are fewer groups than the batch, like ``rd.sample`` vs ``rd.choice`` there),
    one uniformly chosen member ``(rel, value)`` of each group as the positive, and a negative value drawn
    uniformly from ``[0, n_neg)`` and RE-DRAWN until ``(group, rel, negative)`` is not a member of the group
    (the ``while True`` loops at dataset.py:260-262 and :303-305) - so, unlike the edge samplers, a batch
    never contains a false negative and ``false_neg`` is None."""

    def __init__(self, group, rel, value, n_neg, batch_size, n_batch, device=None, generator=None):
        import torch as th
        self.th = th
        self.device = th.device("cpu") if device is None else th.device(device)
        group = np.asarray(group, np.int64)
        value = np.asarray(value, np.int64)
        rel = np.zeros_like(group) if rel is None else np.asarray(rel, np.int64)
        order = np.argsort(group, kind="stable")               # members keep their input order inside a group
        g_sorted = group[order]
        keys, start = np.unique(g_sorted, return_index=True)
        to = lambda a: th.as_tensor(np.ascontiguousarray(a), dtype=th.long, device=self.device)
        self.keys, self.ptr = to(keys), to(np.append(start, len(g_sorted)))
        self.m_rel, self.m_val = to(rel[order]), to(value[order])
        self.n_neg, self.batch_size, self.n_batch = int(n_neg), int(batch_size), int(n_batch)
        self._n_rel = int(rel.max()) + 1 if len(rel) else 1
        self._members = th.unique(self._key(to(g_sorted), self.m_rel, self.m_val))      # sorted
        self.gen = generator

    def _key(self, g, r, v):
        return (g * self._n_rel + r) * self.n_neg + v

    def _is_member(self, g, r, v):
        th = self.th
        inside = (v >= 0) & (v < self.n_neg)           # values outside the negative range cannot collide with a draw
        q = self._key(g, r, v.clamp(0, self.n_neg - 1))
        pos = th.searchsorted(self._members, q).clamp_(max=self._members.shape[0] - 1)
        return inside & (self._members[pos] == q)

    def __iter__(self):
        th = self.th
        G, B = self.keys.shape[0], self.batch_size
        for _ in range(self.n_batch):
            if B <= G:
                gi = th.randperm(G, device=self.device, generator=self.gen)[:B]
            else:
                gi = th.randint(0, G, (B,), device=self.device, generator=self.gen)
            lo, n = self.ptr[gi], self.ptr[gi + 1] - self.ptr[gi]
            pick = lo + (th.rand(B, device=self.device, generator=self.gen) * n).long().minimum(n - 1)
            g, r, v = self.keys[gi], self.m_rel[pick], self.m_val[pick]
            neg = th.randint(0, self.n_neg, (B,), device=self.device, generator=self.gen)
            bad = self._is_member(g, r, neg)
            for _ in range(1000):                       # rejection: only the collided draws are repeated
                idx = bad.nonzero().squeeze(1)
                if idx.numel() == 0:
                    break
                neg[idx] = th.randint(0, self.n_neg, (idx.numel(),), device=self.device, generator=self.gen)
                bad = self._is_member(g, r, neg)
            else:
                raise RuntimeError("negative sampling did not terminate: a group contains every candidate")
            yield g, r, v, neg


class DataLoader(object):
    def __init__(self, data_name, symmetric=True, as_continuous_ids=True, add_uv2KG=True, seed=1234,
                 data_dir=None, _frames=None):
        self._data_name = data_name
        rd.seed(seed)
        self._rng = np.random.default_rng(seed)
        if _frames is None:
            if data_dir is None:
                root = os.environ.get("KGAT_DATA_ROOT", os.path.join(os.getcwd(), "datasets"))
                data_dir = os.path.join(root, data_name, "data")
            train_pd, valid_pd, test_pd = (_read_table(os.path.join(data_dir, f)) for f in _UV_FILES)
            kg_pd = _read_table(os.path.join(data_dir, _KG_FILE))
        else:
            train_pd, valid_pd, test_pd, kg_pd = _frames
        if len(train_pd.columns) != 2:
            raise NotImplementedError("only implicit feedback (columns u, v) is supported, "
                                      "as in the reference's published pipelines")
        train_pd, valid_pd, test_pd = (f.sort_values(by=["u"], kind="stable").reset_index(drop=True)
                                       for f in (train_pd, valid_pd, test_pd))
        self.n_users = int(train_pd["u"].nunique())             # dataset.py:29
        self.n_items = int(train_pd["v"].nunique())             # dataset.py:30
        self.n_train, self.n_valid, self.n_test = train_pd.shape[0], valid_pd.shape[0], test_pd.shape[0]

        tu, tv = train_pd["u"].values, train_pd["v"].values
        vu, vv = valid_pd["u"].values, valid_pd["v"].values
        # dicts hold RAW item ids: built before the offset (dataset.py:45-50)
        self.train_user_dict = _pairs_to_dict(tu, tv)
        self.valid_user_dict = _pairs_to_dict(vu, vv)
        self.test_user_dict = _pairs_to_dict(test_pd["u"].values, test_pd["v"].values)
        self.train_valid_user_dict = _pairs_to_dict(np.concatenate([tu, vu]), np.concatenate([tv, vv]))

        off = self.n_users if as_continuous_ids else 0          # dataset.py:60-73
        self.item_id_range = np.arange(off, off + self.n_items)
        self.train_pairs = np.stack([tu, tv + off], 1).astype(np.int32)
        self.valid_pairs = np.stack([vu, vv + off], 1).astype(np.int32)
        self.test_pairs = np.stack([test_pd["u"].values, test_pd["v"].values + off], 1).astype(np.int32)
        kg = kg_pd[["h", "r", "t"]].values.astype(np.int32).copy()
        kg[:, 0] += off
        kg[:, 2] += off

        n_kg_rel = int(np.unique(kg[:, 1]).size)                # dataset.py:77
        self._train_uv_triplet = self._uv2triplet(self.train_pairs, n_kg_rel, symmetric)
        train_valid_uv = self._uv2triplet(np.vstack([self.train_pairs, self.valid_pairs]), n_kg_rel, symmetric)
        if add_uv2KG:                                            # dataset.py:84-93
            self.train_KG_triplet = np.vstack((kg, self._train_uv_triplet))
            self.test_KG_triplet = np.vstack((kg, train_valid_uv))
        else:
            self.train_KG_triplet = kg
            self.test_KG_triplet = kg
        self.n_KG_relation = int(np.unique(self.train_KG_triplet[:, 1]).size)      # dataset.py:95
        self.n_KG_entity = int(np.unique(np.concatenate(
            (self.train_KG_triplet[:, 0], self.train_KG_triplet[:, 2]))).size)     # dataset.py:96
        self.n_train_KG_triplet = self.train_KG_triplet.shape[0]
        self.n_test_KG_triplet = self.test_KG_triplet.shape[0]
        self._graphs = {}

    # ------------------------------------------------------------------ construction helpers
    @classmethod
    def from_synthetic(cls, config="tiny", seed=1234, skew=False):
        """Same object from a synthetic CKG of a BASELINE.json shape."""
        ck = synth.synth_ckg(config, seed=seed, skew=skew)
        U = ck.cfg.n_users
        frame = lambda p: pd.DataFrame({"u": p[:, 0].astype(np.int32), "v": (p[:, 1] - U).astype(np.int32)})
        kg = pd.DataFrame({"h": ck.kg[:, 0] - U, "r": ck.kg[:, 1], "t": ck.kg[:, 2] - U}, dtype=np.int32)
        half = ck.valid_pairs.shape[0] // 2            # held-out pairs: first half validation, second half test
        return cls("synthetic-" + ck.cfg.name, seed=seed,
                   _frames=(frame(ck.train_pairs), frame(ck.valid_pairs[:half]), frame(ck.valid_pairs[half:]), kg))

    @staticmethod
    def _uv2triplet(pairs, offset_rel, symmetric=True):
        trip = synth.uv_triplets(pairs, offset_rel)              # dataset.py:165-185
        return trip if symmetric else trip[: pairs.shape[0]]

    def _graph(self, name, triplets):
        if name not in self._graphs:
            self._graphs[name] = CKGraph(self.n_KG_entity, triplets, n_rel=self.n_KG_relation)
        return self._graphs[name]

    @property
    def train_g(self):
        return self._graph("train", self.train_KG_triplet)

    @property
    def test_g(self):
        return self._graph("test", self.test_KG_triplet)

    def __repr__(self):
        info = "{}:\n".format(self._data_name)
        info += "\t#user {}, #item {}\n".format(self.n_users, self.n_items)
        info += "\t#train {}, #valid {}, #test {}\n".format(self.n_train, self.n_valid, self.n_test)
        info += "In the knowledge graph:\n"
        info += "\tKG: #entity {}, #relation {}, #train triplet {}, #test triplet {}\n".format(
            self.n_KG_entity, self.n_KG_relation, self.n_train_KG_triplet, self.n_test_KG_triplet)
        return info

    # ------------------------------------------------------------------ samplers (host side)
    def use_device_samplers(self, device):
        """Make KG_sampler / CF_pair_sampler yield DEVICE tensors produced on `device`
        (None switches back to the numpy samplers)."""
        self._sampler_device = device
        self.__dict__.pop("_samplers", None)

    def _sampler(self, name, src, dst, etype, n_nodes, batch_size, mode):
        """Samplers are kept across epochs (their sorted edge-key table is built once)."""
        cache = self.__dict__.setdefault("_samplers", {})
        dev = self.__dict__.get("_sampler_device")
        key = (name, batch_size, mode, str(dev))
        if key not in cache:
            if dev is None:
                cache[key] = EdgeBatchSampler(src, dst, etype, n_nodes, batch_size, mode, self._rng)
            else:
                cache[key] = DeviceEdgeBatchSampler(src, dst, etype, n_nodes, batch_size, mode, dev)
        return cache[key]

    def _unique_sampler(self, name, group, rel, value, n_neg, batch_size, n):
        """dataset.py:236-243 / :285-292: one batch of everything for batch_size < 0 or > n, else n // bs + 1."""
        if batch_size < 0 or batch_size > n:
            batch_size, n_batch = n, 1
        else:
            n_batch = n // batch_size + 1
        cache = self.__dict__.setdefault("_samplers", {})
        dev = self.__dict__.get("_sampler_device")
        key = (name + "-unique", batch_size, str(dev))
        if key not in cache:
            import torch as th
            gen = None
            if dev is None:
                gen = th.Generator().manual_seed(int(self._rng.integers(0, 2 ** 31 - 1)))
            cache[key] = UniqueGroupSampler(group, rel, value, n_neg, batch_size, n_batch, dev, gen)
        return cache[key]

    @staticmethod
    def _n_batch(batch_size, n):
        if batch_size < 0 or batch_size > n:
            return n
        return batch_size

    def KG_sampler(self, batch_size, pos_mode="uniform", neg_mode="tail", num_workers=8, exclude_positive=False):
        """Yields ``(h, r, pos_t, neg_t, false_neg)`` batches (dataset.py:234-282).  As in the
        reference the names refer to the t->h graph: "h" is the edge SOURCE (= KG tail), "pos_t"
        the edge destination (= KG head), and the corrupted node is the destination."""
        t = self.train_KG_triplet
        if pos_mode == "unique" and neg_mode == "tail":
            # dataset.py:246-266: distinct heads, one of the head's (r, t) pairs, a tail that is not one of
            # the head's tails under that relation; n_batch as at dataset.py:236-243
            sampler = self._unique_sampler("kg", t[:, 0], t[:, 1], t[:, 2], self.n_KG_entity, batch_size,
                                           self.n_train_KG_triplet)
            for h, r, pos_t, neg_t in sampler:
                yield h, r, pos_t, neg_t, None
            return
        if not (pos_mode == "uniform" and neg_mode == "tail"):
            raise NotImplementedError("KG_sampler modes: ('uniform', 'tail') and ('unique', 'tail'), as in the reference")
        sampler = self._sampler("kg", t[:, 2], t[:, 0], t[:, 1], self.n_KG_entity,
                                self._n_batch(batch_size, self.n_train_KG_triplet), "tail")
        for s, d, r, _, nd, false_neg in sampler:
            yield s, r, d, nd, false_neg

    def CF_pair_sampler(self, batch_size, pos_mode="uniform", neg_mode="random", num_workers=8,
                        exclude_positive=False):
        """Yields ``(users, pos_items, neg_items, false_neg)`` batches over the symmetric
        user-item graph (dataset.py:283-325): users = edge destinations, pos_items = edge sources,
        neg_items = uniformly corrupted sources."""
        if pos_mode == "unique" and neg_mode == "exclude_pos":
            # dataset.py:294-310: distinct users, one of the user's train items, an item the user has not
            # interacted with.  As in the reference the items are RAW ids in [0, n_items) - train_user_dict is
            # built before the id offset (dataset.py:45-50) - so a caller adds n_users to index the embedding.
            u = np.concatenate([np.full(len(v), k, np.int64) for k, v in self.train_user_dict.items()])
            v = np.concatenate([np.asarray(v, np.int64) for v in self.train_user_dict.values()])
            sampler = self._unique_sampler("cf", u, None, v, self.n_items, batch_size, self.n_train)
            for users, _, pos_items, neg_items in sampler:
                yield users, pos_items, neg_items, None
            return
        if not (pos_mode == "uniform" and neg_mode == "random"):
            raise NotImplementedError("CF_pair_sampler modes: ('uniform', 'random') and ('unique', 'exclude_pos'), "
                                      "as in the reference")
        t = self._train_uv_triplet
        sampler = self._sampler("cf", t[:, 2], t[:, 0], None, self.n_users + self.n_items,
                                self._n_batch(batch_size, t.shape[0]), "head")
        for s, d, _, ns, _, false_neg in sampler:
            yield d, s, ns, false_neg
