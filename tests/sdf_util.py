"""Hand-made MDL V2000 records for the SDF reader tests."""

import numpy as np


def molblock(atoms, bonds, charges=None):
    """atoms: [(symbol, x, y, z)], bonds: [(i, j, order)] 1-based."""
    lines = ["", "  handmade", "", f"{len(atoms):3d}{len(bonds):3d}  0  0  0  0  0  0  0  0999 V2000"]
    for s, x, y, z in atoms:
        lines.append(f"{x:10.4f}{y:10.4f}{z:10.4f} {s:<3s} 0  0  0  0  0  0  0  0  0  0  0  0")
    for i, j, o in bonds:
        lines.append(f"{i:3d}{j:3d}{o:3d}  0")
    if charges:
        lines.append(f"M  CHG{len(charges):3d}" + "".join(f"{a:4d}{c:4d}" for a, c in charges))
    lines.append("M  END")
    return "\n".join(lines) + "\n$$$$\n"


def ring(n, syms, orders, extra_atoms=(), extra_bonds=()):
    atoms = [(syms[i], np.cos(2 * np.pi * i / n) * 1.4, np.sin(2 * np.pi * i / n) * 1.4, 0.0) for i in range(n)]
    bonds = [(i + 1, (i + 1) % n + 1, orders[i]) for i in range(n)]
    atoms += list(extra_atoms)
    bonds += list(extra_bonds)
    return atoms, bonds
