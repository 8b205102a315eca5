"""Synthetic collaborative knowledge graphs of the shapes BASELINE.json names.

There is no network for the real Last-FM / Yelp2018 / Amazon-book dumps, so the
benchmarks and the parity tests use graphs with the same node/edge/relation
counts.  The id layout is the reference's: users ``[0, U)``, items
``[U, U+I)``, remaining KG entities ``[U+I, U+ENT)`` (reference
``dataset.py:60-73``), and the triplet array is ``[KG ; (u, R_kg, v) ;
(v, R_kg+1, u)]`` in int32 ``(h, r, t)`` rows (reference ``dataset.py:89``,
``dataset.py:165-185``).  SURVEY.md section 8(d) fixes the generator.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

__all__ = ["CKGConfig", "CONFIGS", "SynthCKG", "synth_ckg"]


@dataclass(frozen=True)
class CKGConfig:
    name: str
    n_users: int
    n_entities: int      # KG entities, items included
    n_items: int
    n_kg_relations: int
    n_kg_triples: int
    n_train: int         # train interaction pairs
    n_valid: int         # held-out pairs used as the eval positives

    @property
    def n_nodes(self) -> int:
        return self.n_users + self.n_entities

    @property
    def n_relations(self) -> int:
        return self.n_kg_relations + 2

    @property
    def n_edges(self) -> int:
        return self.n_kg_triples + 2 * self.n_train


CONFIGS = {
    c.name: c
    for c in (
        CKGConfig("tiny", 2_000, 5_000, 2_000, 9, 40_000, 20_000, 4_000),
        CKGConfig("last-fm", 23_566, 58_266, 48_123, 9, 464_567, 2_185_053, 242_784),
        CKGConfig("yelp2018", 45_919, 90_961, 45_538, 42, 1_853_704, 819_310, 110_722),
        CKGConfig("amazon-book", 70_679, 88_572, 24_915, 39, 2_557_746, 552_778, 99_736),
        CKGConfig("amazon-book-10x", 706_790, 1_000_000, 249_150, 39, 25_577_460, 5_527_780, 997_360),
    )
}


@dataclass
class SynthCKG:
    cfg: CKGConfig
    kg: np.ndarray             # (n_kg_triples, 3) int32 (h, r, t), ids already offset by n_users
    train_pairs: np.ndarray    # (n_train, 2) int32 (u, v) with v offset by n_users
    valid_pairs: np.ndarray    # (<= n_valid, 2) int32, de-duplicated against train per user
    triplets: np.ndarray       # (E, 3) int32 train CKG, reference row order

    @property
    def item_id_range(self) -> np.ndarray:
        u = self.cfg.n_users
        return np.arange(u, u + self.cfg.n_items)


def uv_triplets(pairs: np.ndarray, offset_rel: int) -> np.ndarray:
    """(u, offset_rel, v) rows followed by (v, offset_rel + 1, u) rows.

    Same layout as reference ``dataset.py:165-185`` with ``n_uv_rel = 1`` and
    ``symmetric=True``.
    """
    n = pairs.shape[0]
    out = np.empty((2 * n, 3), dtype=np.int32)
    out[:n, 0] = pairs[:, 0]
    out[:n, 1] = offset_rel
    out[:n, 2] = pairs[:, 1]
    out[n:, 0] = pairs[:, 1]
    out[n:, 1] = offset_rel + 1
    out[n:, 2] = pairs[:, 0]
    return out


def _draw_pairs(rng, cfg: CKGConfig, n: int, skew: bool, cover: bool) -> np.ndarray:
    U, I = cfg.n_users, cfg.n_items
    u = rng.integers(0, U, size=n, dtype=np.int64)
    if skew:
        v = rng.zipf(1.5, size=n) % I
    else:
        v = rng.integers(0, I, size=n, dtype=np.int64)
    if cover:
        # every user and every item occurs in train (reference counts both by
        # nunique over the train split, dataset.py:29-30)
        u[: min(U, n)] = np.arange(min(U, n))
        v[: min(I, n)] = np.arange(min(I, n))
    return np.stack([u, v + U], axis=1).astype(np.int32)


def synth_ckg(cfg: CKGConfig | str, seed: int = 1234, skew: bool = False) -> SynthCKG:
    if isinstance(cfg, str):
        cfg = CONFIGS[cfg]
    rng = np.random.default_rng(seed)
    U, ENT, Rk, T = cfg.n_users, cfg.n_entities, cfg.n_kg_relations, cfg.n_kg_triples
    h = rng.integers(0, ENT, size=T, dtype=np.int64)
    r = rng.integers(0, Rk, size=T, dtype=np.int64)
    if skew:
        t = rng.zipf(1.5, size=T) % ENT
    else:
        t = rng.integers(0, ENT, size=T, dtype=np.int64)
    # contiguity: every entity id and every relation id is present, so the
    # reference's "#unique" counts (dataset.py:95-96) equal the id ranges
    t[: min(ENT, T)] = np.arange(min(ENT, T))
    r[: min(Rk, T)] = np.arange(min(Rk, T))
    kg = np.stack([h + U, r, t + U], axis=1).astype(np.int32)

    train = _draw_pairs(rng, cfg, cfg.n_train, skew, cover=True)
    valid = _draw_pairs(rng, cfg, cfg.n_valid, skew, cover=False)
    # drop held-out pairs that are already train pairs of the same user
    I = cfg.n_items
    key_t = train[:, 0].astype(np.int64) * I + (train[:, 1] - U)
    key_v = valid[:, 0].astype(np.int64) * I + (valid[:, 1] - U)
    keep = ~np.isin(key_v, key_t)
    _, first = np.unique(key_v, return_index=True)
    uniq = np.zeros(len(key_v), dtype=bool)
    uniq[first] = True
    valid = valid[keep & uniq]

    trip = np.vstack([kg, uv_triplets(train, Rk)])
    return SynthCKG(cfg, kg, train, valid, trip)


def pairs_to_user_dict(pairs: np.ndarray, n_users: int) -> dict:
    """{user: np.array(raw item ids)} in ascending user order (dataset.py:187-196)."""
    order = np.argsort(pairs[:, 0], kind="stable")
    u = pairs[order, 0]
    v = pairs[order, 1] - n_users
    uniq, start = np.unique(u, return_index=True)
    bounds = np.append(start, len(u))
    return {int(uu): v[bounds[i]: bounds[i + 1]] for i, uu in enumerate(uniq)}
