"""Seeded synthetic inputs for the scoring path: typed ligand topologies and conformer coordinates.

There is no chemistry toolkit in the build or benchmark environment, so benchmark
and parity inputs are *typed ligand graphs* emitted directly: 3-6 fragments per ligand drawn from
{aromatic ring, H-bond group, carboxylate, hydrophobic chain, tertiary amine, halogen}, fragment centres on a
random walk, conformers = rigid fragments displaced by a smooth per-conformer bend plus atom noise
(SURVEY.md section 8d). The same `TypedLigand` objects feed this package's host featuriser and, in the
golden-vector script, the reference's own `LigandGraph`.

Nothing here is on the product hot path; it only creates data.
"""

from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from .ligand import TypedLigand

_RING_R = 1.39  # A, benzene-like hexagon radius
_BOND = 1.5


@dataclass
class LigandTemplate:
    """Topology + rigid local geometry of one synthetic ligand."""

    atomic_nums: list[int]
    neighbors: list[list[int]]
    pharmacophores: list[tuple[str, int | tuple[int, ...], int | tuple[int, ...]]]
    frag_of_atom: np.ndarray  # int [N]
    local_xyz: np.ndarray  # float32 [N, 3] offset of the atom from its fragment centre
    n_frag: int

    def typed(self, atom_positions: np.ndarray | None = None, name: str = "") -> TypedLigand:
        return TypedLigand(self.atomic_nums, self.neighbors, self.pharmacophores, atom_positions, name)


class _Builder:
    def __init__(self):
        self.z: list[int] = []
        self.nb: list[list[int]] = []
        self.frag: list[int] = []
        self.xyz: list[tuple[float, float, float]] = []

    def atom(self, z: int, frag: int, xyz) -> int:
        self.z.append(z)
        self.nb.append([])
        self.frag.append(frag)
        self.xyz.append(tuple(float(v) for v in xyz))
        return len(self.z) - 1

    def bond(self, a: int, b: int) -> None:
        self.nb[a].append(b)
        self.nb[b].append(a)


def make_template(rng: np.random.Generator, frag_range: tuple[int, int] = (3, 7)) -> LigandTemplate:
    """Draw one ligand topology. Typing follows the *definitions* in ligand_utils.py:36-88 on the toy graph:
    hydrophobic = carbon with only carbon neighbours; acceptors = ring N, hydroxyl/carbonyl O, amine N;
    donors = hydroxyl O / primary amine N; cation = tertiary amine N; anion = carboxylate; halogen = C-X."""
    b = _Builder()
    n_frag = int(rng.integers(frag_range[0], frag_range[1]))
    rings: list[tuple[int, ...]] = []
    cations: list[int] = []
    anions: list[tuple[int, tuple[int, ...], tuple[int, ...]]] = []
    donors: list[int] = []
    acceptors: list[int] = []
    halogens: list[int] = []
    anchors: list[int] = []  # atom of each fragment that bonds to the next fragment
    kinds = rng.choice(
        ["ring", "hbond", "carboxylate", "chain", "amine", "halogen"],
        size=n_frag,
        p=[0.34, 0.2, 0.08, 0.2, 0.1, 0.08],
    )
    if "ring" not in kinds and rng.random() < 0.7:
        kinds[int(rng.integers(0, n_frag))] = "ring"
    for f, kind in enumerate(kinds):
        prev = anchors[-1] if anchors else None
        if kind == "ring":
            ids = []
            n_pos = int(rng.integers(0, 6)) if rng.random() < 0.35 else -1
            for k in range(6):
                ang = np.pi / 3 * k
                z = 7 if k == n_pos else 6
                ids.append(b.atom(z, f, (_RING_R * np.cos(ang), _RING_R * np.sin(ang), 0.0)))
            for k in range(6):
                b.bond(ids[k], ids[(k + 1) % 6])
            rings.append(tuple(sorted(ids)))
            if n_pos >= 0:
                acceptors.append(ids[n_pos])
            attach = ids[(n_pos + 3) % 6 if n_pos >= 0 else 0]
            if prev is not None:
                b.bond(prev, attach)
            anchors.append(ids[(ids.index(attach) + 3) % 6] if b.z[ids[(ids.index(attach) + 3) % 6]] == 6 else attach)
        elif kind == "hbond":
            c = b.atom(6, f, (0.0, 0.0, 0.0))
            o = b.atom(8, f, (1.2, 0.3, 0.0))
            b.bond(c, o)
            acceptors.append(o)
            if rng.random() < 0.5:
                donors.append(o)
            if rng.random() < 0.4:  # amide-like second hetero atom on the same carbon
                n2 = b.atom(7, f, (-0.7, 1.1, 0.2))
                b.bond(c, n2)
                donors.append(n2)
            if prev is not None:
                b.bond(prev, c)
            anchors.append(c)
        elif kind == "carboxylate":
            c = b.atom(6, f, (0.0, 0.0, 0.0))
            o1 = b.atom(8, f, (1.1, 0.6, 0.0))
            o2 = b.atom(8, f, (-1.1, 0.6, 0.0))
            b.bond(c, o1)
            b.bond(c, o2)
            acceptors += [o1, o2]
            anions.append((c, (c, o1, o2), (o1, o2)))
            if prev is not None:
                b.bond(prev, c)
            anchors.append(c)
        elif kind == "chain":
            n_c = int(rng.integers(1, 4))
            last = prev
            first = None
            for k in range(n_c):
                a = b.atom(6, f, (_BOND * 0.85 * (k - (n_c - 1) / 2), 0.5 * (k % 2), 0.0))
                if last is not None:
                    b.bond(last, a)
                last = a
                first = a if first is None else first
            if n_c >= 2 and rng.random() < 0.3:  # branch methyl
                m = b.atom(6, f, (0.0, -1.3, 0.6))
                b.bond(first, m)
            anchors.append(last)
        elif kind == "amine":
            n = b.atom(7, f, (0.0, 0.0, 0.0))
            m1 = b.atom(6, f, (1.2, 0.7, 0.3))
            m2 = b.atom(6, f, (-1.2, 0.7, -0.3))
            b.bond(n, m1)
            b.bond(n, m2)
            cations.append(n)
            acceptors.append(n)
            if prev is not None:
                b.bond(prev, n)
            anchors.append(n)
        elif kind == "halogen":
            c = b.atom(6, f, (0.0, 0.0, 0.0))
            x = b.atom(int(rng.choice([9, 17, 35])), f, (1.7, 0.0, 0.0))
            b.bond(c, x)
            halogens.append(x)
            if prev is not None:
                b.bond(prev, c)
            anchors.append(c)
    hydrophobics = [
        i for i, z in enumerate(b.z) if z == 6 and all(b.z[j] in (1, 6) for j in b.nb[i])
    ]
    ph: list[tuple[str, int | tuple[int, ...], int | tuple[int, ...]]] = []
    ph += [("Hydrophobic", i, i) for i in hydrophobics]
    ph += [("Aromatic", r, r) for r in sorted(rings)]
    ph += [("Cation", i, i) for i in cations]
    ph += [("Anion", grp, ctr) for _, grp, ctr in anions]
    ph += [("HBond_donor", i, i) for i in sorted(donors)]
    ph += [("HBond_acceptor", i, i) for i in sorted(acceptors)]
    ph += [("Halogen", i, i) for i in halogens]
    return LigandTemplate(
        atomic_nums=b.z,
        neighbors=b.nb,
        pharmacophores=ph,
        frag_of_atom=np.asarray(b.frag, dtype=np.int64),
        local_xyz=np.asarray(b.xyz, dtype=np.float32),
        n_frag=n_frag,
    )


def _random_rotations(rng: np.random.Generator, n: int) -> np.ndarray:
    q = rng.normal(size=(n, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    w, x, y, z = q.T
    return np.stack(
        [
            1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w),
            2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
            2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y),
        ],
        axis=1,
    ).reshape(n, 3, 3)


def make_conformers(
    tmpl: LigandTemplate,
    num_conformers: int,
    rng: np.random.Generator,
    flex: float = 0.45,
    noise: float = 0.08,
) -> np.ndarray:
    """float32 [N_atoms, C, 3]: fragment centres on a 2.5-4 A random walk, one rigid rotation per fragment,
    per-conformer cumulative displacement of the centres (std `flex` A per step) and atom noise."""
    F = tmpl.n_frag
    steps = rng.normal(size=(F, 3))
    steps /= np.linalg.norm(steps, axis=1, keepdims=True)
    steps *= rng.uniform(2.5, 4.0, size=(F, 1))
    steps[0] = 0.0
    centres = np.cumsum(steps, axis=0)  # [F,3]
    rot = _random_rotations(rng, F)  # [F,3,3]
    bend = np.cumsum(rng.normal(scale=flex, size=(num_conformers, F, 3)), axis=1)  # [C,F,3]
    bend[:, 0] = 0.0
    f = tmpl.frag_of_atom
    local = np.einsum("nij,nj->ni", rot[f], tmpl.local_xyz.astype(np.float64))  # [N,3]
    pos = centres[f][:, None, :] + bend[:, f, :].transpose(1, 0, 2) + local[:, None, :]
    pos = pos + rng.normal(scale=noise, size=pos.shape)
    return np.ascontiguousarray(pos.astype(np.float32))


def make_ligands(
    n: int,
    num_conformers: int,
    seed: int,
    flex: float = 0.45,
    noise: float = 0.08,
    frag_range: tuple[int, int] = (3, 7),
) -> list[TypedLigand]:
    """n independent synthetic typed ligands with coordinates (host, numpy). Deterministic in `seed`."""
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n):
        t = make_template(rng, frag_range)
        while len(t.pharmacophores) == 0:
            t = make_template(rng, frag_range)
        out.append(t.typed(make_conformers(t, num_conformers, rng, flex, noise), name=f"syn{seed}_{i}"))
    return out


def make_hotspot_infos(
    seed: int = 0, n_hotspots: int = 40, r_lo: float = 0.9, r_hi: float = 1.6, size: int = 64, type_probs=None
):
    """Synthetic '6OIM-like' hotspot list for `PharmacophoreModel.create`: Gaussian blobs
    on the 64^3 grid, zeroed below 0.5, one NCI type each."""
    from .constants import INTERACTION_LIST

    rng = np.random.default_rng(seed)
    p = np.array([0.35, 0.05, 0.05, 0.03, 0.03, 0.17, 0.17, 0.05, 0.05, 0.05] if type_probs is None else type_probs, dtype=np.float64)
    g = np.arange(size, dtype=np.float64)
    infos = []
    for _ in range(n_hotspots):
        typ = INTERACTION_LIST[int(rng.choice(len(INTERACTION_LIST), p=p / p.sum()))]
        c = rng.normal(size / 2, 6.0, size=3)
        r = rng.uniform(r_lo, r_hi)
        d2 = (
            (g[:, None, None] - c[0]) ** 2 + (g[None, :, None] - c[1]) ** 2 + (g[None, None, :] - c[2]) ** 2
        )
        m = np.exp(-0.5 * d2 / (r * r))
        m[m < 0.5] = 0.0
        infos.append(
            dict(
                nci_type=typ,
                hotspot_position=np.asarray(c * 0.5 - (size - 1) * 0.25, dtype=np.float64),
                hotspot_score=0.9,
                point_map=m,
            )
        )
    return infos


def make_library_device(
    n_ligands: int,
    num_conformers: int,
    seed: int,
    device="cuda",
    n_templates: int = 4096,
    flex: float = 0.45,
    noise: float = 0.08,
    keep_atoms: int = 0,
    coord_seed: int | None = None,
):
    """Large synthetic library built on the GPU (benchmarks): `n_templates` topologies from `make_template`, each
    replicated with independently drawn coordinates (same recipe as `make_conformers`, in torch on `device`).
    `seed` draws the topologies, `coord_seed` (default: `seed`) the coordinates: the shards of several ranks share
    the topology set and differ in their conformers, so that every rank holds the same amount of work (a random
    shard of a real library has the same cost distribution as any other; 4096 topologies drawn per rank do not).
    Ligand i uses template i % n_templates, so any prefix of the library is a representative sample.
    Returns a `scoring.DeviceLigandBatch`. Data creation only - nothing here is scored or timed.
    keep_atoms = n (<= n_templates): also keep the ATOM coordinates of the first n ligands and attach them as
    `batch.typed_prefix` (list of TypedLigand) - the same ligands in the form the reference's own `LigandGraph`
    consumes (bench.py's reference arm)."""
    import torch

    from .ligand import build_topology
    from .scoring import DeviceLigandBatch

    dev = torch.device(device)
    rng = np.random.default_rng(seed)
    T = max(1, min(n_templates, n_ligands))
    C = int(num_conformers)
    stride = (C + 3) // 4 * 4
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed if coord_seed is None else coord_seed)

    keep_atoms = min(int(keep_atoms), T)
    typed_prefix: list = []
    templates, tops = [], []
    for _ in range(T):
        t = make_template(rng)
        while len(t.pharmacophores) == 0:
            t = make_template(rng)
        templates.append(t)
        tops.append(build_topology(t.typed()))
    ordered = [top.ordered_cluster_nodes() for top in tops]
    tid = np.arange(n_ligands, dtype=np.int64) % T

    def offsets(per_ligand: np.ndarray) -> np.ndarray:
        off = np.zeros(len(per_ligand) + 1, dtype=np.int64)
        np.cumsum(per_ligand, out=off[1:])
        return off

    nn_t = np.asarray([top.num_nodes for top in tops], dtype=np.int64)
    ncl_t = np.asarray([len(o) for o in ordered], dtype=np.int64)
    ncn_t = np.asarray([sum(len(c) for c in o) for o in ordered], dtype=np.int64)
    node_off = offsets(nn_t[tid])
    clu_off = offsets(ncl_t[tid])
    cn_lig_off = offsets(ncn_t[tid])  # first cluster-node of each ligand
    coord_off = offsets(nn_t[tid] * 3 * stride)
    assert node_off[-1] < 2**31 and cn_lig_off[-1] < 2**31

    masks = np.empty(int(node_off[-1]), dtype=np.uint8)
    cl_nodes = np.empty(int(cn_lig_off[-1]), dtype=np.uint8)
    cn_off = np.empty(int(clu_off[-1]) + 1, dtype=np.int64)
    cn_off[-1] = cn_lig_off[-1]
    coords = torch.empty(int(coord_off[-1]), dtype=torch.float32, device=dev)
    for t_i, (t, top) in enumerate(zip(templates, tops)):
        ligs = np.arange(t_i, n_ligands, T)
        r = len(ligs)
        if r == 0:
            continue
        nn = int(nn_t[t_i])
        masks[(node_off[ligs][:, None] + np.arange(nn)[None, :]).reshape(-1)] = np.tile(top.node_type_mask, r)
        flat = np.asarray([n for c in ordered[t_i] for n in c], dtype=np.uint8)
        cl_nodes[(cn_lig_off[ligs][:, None] + np.arange(len(flat))[None, :]).reshape(-1)] = np.tile(flat, r)
        starts = np.concatenate([[0], np.cumsum([len(c) for c in ordered[t_i]])[:-1]]).astype(np.int64)
        cn_off[(clu_off[ligs][:, None] + np.arange(len(starts))[None, :]).reshape(-1)] = (
            cn_lig_off[ligs][:, None] + starts[None, :]
        ).reshape(-1)
        # ---- coordinates on the device: [r, Nn, 3, stride]
        F, N = t.n_frag, len(t.atomic_nums)
        f = torch.as_tensor(t.frag_of_atom, device=dev)
        steps = torch.randn((r, F, 3), generator=gen, device=dev, dtype=torch.float32)
        steps = steps / steps.norm(dim=-1, keepdim=True) * (2.5 + 1.5 * torch.rand((r, F, 1), generator=gen, device=dev))
        steps[:, 0] = 0.0
        centres = steps.cumsum(dim=1)
        q = torch.randn((r, F, 4), generator=gen, device=dev, dtype=torch.float32)
        q = q / q.norm(dim=-1, keepdim=True)
        w, x, y, z = q.unbind(-1)
        rot = torch.stack(
            [
                1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w),
                2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
                2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y),
            ],
            dim=-1,
        ).reshape(r, F, 3, 3)  # fmt: skip
        bend = (flex * torch.randn((r, C, F, 3), generator=gen, device=dev, dtype=torch.float32)).cumsum(dim=2)
        bend[:, :, 0] = 0.0
        local = torch.einsum("rnij,nj->rni", rot[:, f], torch.as_tensor(t.local_xyz, device=dev))
        pos = centres[:, f][:, :, None, :] + bend[:, :, f, :].permute(0, 2, 1, 3) + local[:, :, None, :]
        pos = pos + noise * torch.randn(pos.shape, generator=gen, device=dev, dtype=torch.float32)  # [r, N, C, 3]
        if t_i < keep_atoms:  # ligand t_i is replica 0 of template t_i
            typed_prefix.append(t.typed(np.ascontiguousarray(pos[0].cpu().numpy()), name=f"lib{seed}_{t_i}"))
        A = torch.zeros((nn, N), dtype=torch.float32, device=dev)
        for n_i, ctr in enumerate(top.node_center_atoms):
            A[n_i, list(ctr)] = 1.0 / len(ctr)
        npos = torch.einsum("na,racx->rnxc", A, pos)  # [r, Nn, 3, C]
        if stride != C:
            npos = torch.nn.functional.pad(npos, (0, stride - C))
        row = nn * 3 * stride
        dst = torch.as_tensor(coord_off[ligs], device=dev)[:, None] + torch.arange(row, device=dev)[None, :]
        coords[dst.reshape(-1)] = npos.reshape(-1)
    host = dict(
        lig_node_off=node_off.astype(np.int32),
        lig_cluster_off=clu_off.astype(np.int32),
        cluster_node_off=cn_off.astype(np.int32),
        cluster_nodes=cl_nodes,
        node_type_mask=masks,
        n_conf=np.full(n_ligands, C, dtype=np.int32),
        coord_off=coord_off,
    )
    tensors = {k: torch.from_numpy(v).to(dev) for k, v in host.items()}
    tensors["coords"] = coords
    if typed_prefix:
        # node coordinates of the kept ligands through the host featuriser (ligand.node_positions: the reference's own
        # fp32 arithmetic, ligand.py:293-301) instead of the einsum above, so that the device library and the
        # reference's LigandGraph built from `typed_prefix` hold bit-identical node positions
        from .packing import LigandBatch

        lb = LigandBatch.from_typed(typed_prefix)
        n_keep = len(typed_prefix)
        assert np.array_equal(lb.coord_off, coord_off[: n_keep + 1]) and np.array_equal(lb.node_type_mask, masks[: int(node_off[n_keep])])
        coords[: int(coord_off[n_keep])] = torch.from_numpy(lb.coords).to(dev)
    batch = DeviceLigandBatch(tensors, n_ligands, n_ligands * C, max_conformers=C)
    batch.typed_prefix = typed_prefix
    return batch
