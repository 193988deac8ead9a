"""SMILES -> molecular graph with the atom / bond properties the reference tokenises, without RDKit
(SURVEY.md 8 f2).

The reference (`graphchem/preprocessing/features.py:66-83, 112-119, 294-319`) turns every atom and bond
into `str()` of a tuple of RDKit properties and numbers the distinct strings in first-seen order.  RDKit
(`rdkit-pypi==2022.9.5`, `pyproject.toml:18`) is not vendored and cannot be installed offline, so this
module restates the published RDKit perception rules those properties depend on, for the chemistry of the
bundled fuel data sets (C/H/O/N/halogens/S/P, rings, aromatics, charges in brackets, @/@@, / and \\):

* SMILES grammar: atoms in written order; chain bonds are created as they are read, ring-closure bonds
  after the whole string, ordered by ring-closure number and then by occurrence (RDKit `CloseMolRings`);
  `atom.GetBonds()` = creation order.  This fixes the edge order of `MoleculeEncoder.encode`.
* valence model (explicit / implicit valence, implicit hydrogens, radicals), ring membership and smallest
  ring size, RDKit's default aromaticity model (electron-donor types, Hueckel 4n+2 on single rings and
  fused combinations of up to six rings), conjugation (`markConjAtomBonds`), hybridisation
  (`numBondsPlusLonePairs`), legacy stereo perception (CIP-style iterative ranks, chiral tags adjusted for
  ring-closure bond order, tags dropped on non-stereocentres, E/Z of double bonds outside rings < 8).

Objects mimic the slice of the RDKit API the reference calls (`GetAtoms`, `GetBonds`, `GetChiralTag`,
`IsInRingSize`, ...), and the enums print like RDKit's Boost.Python enums, so `atom_to_str` / `bond_to_str`
produce the very strings pinned in the reference's tests (`tests/test_features.py:43-88`).

What pins it (tests/test_smiles.py): those strings, the four notebook encodings and the
eight vocabulary sizes the notebooks print for the CN / YSI / RON / LHV training splits.
"""
from __future__ import annotations

import itertools
from typing import Dict, List, Optional, Tuple


class _Enum:
    """Prints like a Boost.Python enum value: repr 'rdkit.Chem.rdchem.<Type>.<NAME>', str '<NAME>'."""
    __slots__ = ("kind", "name")

    def __init__(self, kind: str, name: str):
        self.kind, self.name = kind, name

    def __repr__(self) -> str:
        return f"rdkit.Chem.rdchem.{self.kind}.{self.name}"

    def __str__(self) -> str:
        return self.name

    def __eq__(self, other) -> bool:
        return isinstance(other, _Enum) and (self.kind, self.name) == (other.kind, other.name)

    def __hash__(self) -> int:
        return hash((self.kind, self.name))

    def __reduce__(self):
        return (_Enum, (self.kind, self.name))


class ChiralType:
    CHI_UNSPECIFIED = _Enum("ChiralType", "CHI_UNSPECIFIED")
    CHI_TETRAHEDRAL_CW = _Enum("ChiralType", "CHI_TETRAHEDRAL_CW")
    CHI_TETRAHEDRAL_CCW = _Enum("ChiralType", "CHI_TETRAHEDRAL_CCW")


class HybridizationType:
    UNSPECIFIED = _Enum("HybridizationType", "UNSPECIFIED")
    S = _Enum("HybridizationType", "S")
    SP = _Enum("HybridizationType", "SP")
    SP2 = _Enum("HybridizationType", "SP2")
    SP3 = _Enum("HybridizationType", "SP3")
    SP3D = _Enum("HybridizationType", "SP3D")
    SP3D2 = _Enum("HybridizationType", "SP3D2")


class BondType:
    SINGLE = _Enum("BondType", "SINGLE")
    DOUBLE = _Enum("BondType", "DOUBLE")
    TRIPLE = _Enum("BondType", "TRIPLE")
    QUADRUPLE = _Enum("BondType", "QUADRUPLE")
    AROMATIC = _Enum("BondType", "AROMATIC")


class BondStereo:
    STEREONONE = _Enum("BondStereo", "STEREONONE")
    STEREOZ = _Enum("BondStereo", "STEREOZ")
    STEREOE = _Enum("BondStereo", "STEREOE")


# symbol -> (atomic number, outer-shell electrons, allowed valences; the first one is the default valence)
_ELEMENTS: Dict[str, Tuple[int, int, Tuple[int, ...]]] = {
    "H": (1, 1, (1,)), "B": (5, 3, (3,)), "C": (6, 4, (4,)), "N": (7, 5, (3,)), "O": (8, 6, (2,)),
    "F": (9, 7, (1,)), "Si": (14, 4, (4,)), "P": (15, 5, (3, 5, 7)), "S": (16, 6, (2, 4, 6)),
    "Cl": (17, 7, (1,)), "Se": (34, 6, (2, 4, 6)), "Br": (35, 7, (1,)), "I": (53, 7, (1, 3, 5)),
}
_BY_NUMBER = {v[0]: v for v in _ELEMENTS.values()}
_ORGANIC = ("Cl", "Br", "B", "C", "N", "O", "P", "S", "F", "I")
_AROMATIC_ORGANIC = {"b": "B", "c": "C", "n": "N", "o": "O", "p": "P", "s": "S"}
_ORDER = {"-": 1.0, "=": 2.0, "#": 3.0, "$": 4.0, ":": 1.5}
_TYPE_OF = {1.0: BondType.SINGLE, 2.0: BondType.DOUBLE, 3.0: BondType.TRIPLE, 4.0: BondType.QUADRUPLE,
            1.5: BondType.AROMATIC}


class Atom:
    def __init__(self, mol: "Mol", idx: int, symbol: str, aromatic: bool = False, charge: int = 0,
                 explicit_hs: int = 0, no_implicit: bool = False, chiral: str = "", isotope: int = 0):
        self._mol, self._idx, self._symbol = mol, idx, symbol
        self._z, self._nouter, self._valences = _ELEMENTS[symbol]
        self._aromatic = aromatic            # as written (lower case); re-perceived by _perceive()
        self._charge, self._explicit_hs, self._no_implicit = charge, explicit_hs, no_implicit
        self._chiral_mark, self._isotope = chiral, isotope
        self._bonds: List["Bond"] = []       # creation order
        self._ring_closures: List[object] = []   # bonds closing rings at this atom, in the order of their digits
        self._smiles_start = False
        self._explicit_valence = 0
        self._implicit_hs = 0
        self._radicals = 0
        self._ring_size = 0
        self._hyb = HybridizationType.UNSPECIFIED
        self._tag = ChiralType.CHI_UNSPECIFIED

    # ---- the RDKit calls of features.py:66-83 and :302-317 ----
    def GetIdx(self) -> int: return self._idx
    def GetSymbol(self) -> str: return self._symbol
    def GetAtomicNum(self) -> int: return self._z
    def GetBonds(self): return tuple(self._bonds)
    def GetChiralTag(self): return self._tag
    def GetDegree(self) -> int: return len(self._bonds)
    def GetExplicitValence(self) -> int: return self._explicit_valence
    def GetFormalCharge(self) -> int: return self._charge
    def GetHybridization(self): return self._hyb
    def GetImplicitValence(self) -> int: return self._implicit_hs
    def GetIsAromatic(self) -> bool: return self._aromatic
    def GetNoImplicit(self) -> bool: return self._no_implicit
    def GetNumExplicitHs(self) -> int: return self._explicit_hs
    def GetNumImplicitHs(self) -> int: return self._implicit_hs
    def GetNumRadicalElectrons(self) -> int: return self._radicals
    def GetTotalDegree(self) -> int: return len(self._bonds) + self.GetTotalNumHs()
    def GetTotalNumHs(self) -> int: return self._explicit_hs + self._implicit_hs
    def GetTotalValence(self) -> int: return self._explicit_valence + self._implicit_hs
    def IsInRing(self) -> bool: return self._ring_size > 0
    def IsInRingSize(self, size: int) -> bool: return size in self._ring_sizes()

    def _ring_sizes(self):
        return {len(r) for r in self._mol._rings if self._idx in r}

    def _neighbors(self):
        return [b._other(self) for b in self._bonds]


class Bond:
    def __init__(self, mol: "Mol", idx: int, a: Atom, b: Atom, order: float, direction: str = ""):
        self._mol, self._idx, self._a, self._b = mol, idx, a, b
        self._order = order                  # Kekule order during perception, 1.5 once perceived aromatic
        self._dir = direction                # '/' or '\\' as written from begin atom to end atom
        self._aromatic = False
        self._conjugated = False
        self._stereo = BondStereo.STEREONONE
        self._in_ring = False

    def GetIdx(self) -> int: return self._idx
    def GetBeginAtom(self) -> Atom: return self._a
    def GetEndAtom(self) -> Atom: return self._b
    def GetBeginAtomIdx(self) -> int: return self._a._idx
    def GetEndAtomIdx(self) -> int: return self._b._idx
    def GetBondType(self): return _TYPE_OF[self._order]
    def GetBondTypeAsDouble(self) -> float: return self._order
    def GetIsAromatic(self) -> bool: return self._aromatic
    def GetIsConjugated(self) -> bool: return self._conjugated
    def GetStereo(self): return self._stereo
    def IsInRing(self) -> bool: return self._in_ring
    def IsInRingSize(self, size: int) -> bool:
        return size in {len(r) for r in self._mol._bond_rings if self._idx in r}

    def _other(self, atom: Atom) -> Atom:
        return self._b if atom is self._a else self._a


class Mol:
    def __init__(self):
        self._atoms: List[Atom] = []
        self._bonds: List[Bond] = []
        self._rings: List[Tuple[int, ...]] = []        # atom index tuples
        self._bond_rings: List[frozenset] = []         # bond index sets, parallel to _rings

    def GetAtoms(self): return tuple(self._atoms)
    def GetBonds(self): return tuple(self._bonds)
    def GetNumAtoms(self) -> int: return len(self._atoms)
    def GetNumBonds(self) -> int: return len(self._bonds)
    def GetAtomWithIdx(self, i: int) -> Atom: return self._atoms[i]
    def GetBondWithIdx(self, i: int) -> Bond: return self._bonds[i]

    def GetBondBetweenAtoms(self, i: int, j: int) -> Optional[Bond]:
        for b in self._atoms[i]._bonds:
            if b._other(self._atoms[i])._idx == j:
                return b
        return None

    def _add_bond(self, a: Atom, b: Atom, order: float, direction: str = "") -> Bond:
        bond = Bond(self, len(self._bonds), a, b, order, direction)
        self._bonds.append(bond)
        a._bonds.append(bond)
        b._bonds.append(bond)
        return bond


class SmilesError(ValueError):
    pass


# ---------------------------------------------------------------- parsing ----

def _parse_bracket(mol: Mol, body: str) -> Atom:
    i, n = 0, len(body)
    iso = 0
    while i < n and body[i].isdigit():
        iso = iso * 10 + int(body[i]); i += 1
    sym, aromatic = None, False
    if body[i:i + 2] in _ELEMENTS and body[i:i + 2] not in ("H",):
        sym = body[i:i + 2]; i += 2
    elif body[i:i + 1] in _ELEMENTS:
        sym = body[i]; i += 1
    elif body[i:i + 2] == "se":
        sym, aromatic = "Se", True; i += 2
    elif body[i:i + 1] in _AROMATIC_ORGANIC:
        sym, aromatic = _AROMATIC_ORGANIC[body[i]], True; i += 1
    if sym is None:
        raise SmilesError(f"unsupported atom [{body}]")
    chiral = ""
    if body[i:i + 2] == "@@":
        chiral = "@@"; i += 2
    elif body[i:i + 1] == "@":
        chiral = "@"; i += 1
    hs = 0
    if body[i:i + 1] == "H":
        i += 1
        hs = 1
        if i < n and body[i].isdigit():
            hs = int(body[i]); i += 1
    charge = 0
    while i < n and body[i] in "+-":
        sign = 1 if body[i] == "+" else -1
        i += 1
        if i < n and body[i].isdigit():
            charge += sign * int(body[i]); i += 1
        else:
            charge += sign
    if i < n and body[i] == ":":            # atom class: ignored
        i += 1
        while i < n and body[i].isdigit():
            i += 1
    if i != n:
        raise SmilesError(f"cannot parse atom [{body}]")
    return Atom(mol, len(mol._atoms), sym, aromatic, charge, hs, True, chiral, iso)


def _parse(smiles: str) -> Mol:
    mol = Mol()
    s = smiles.strip()
    if not s:
        raise SmilesError("empty SMILES")
    s = s.split()[0]                        # RDKit reads "SMILES name": everything after white space is a title
    prev: Optional[Atom] = None
    stack: List[Optional[Atom]] = []
    pending = ""                            # bond symbol waiting for its second atom
    open_rings: Dict[int, List[Tuple[Atom, str]]] = {}   # ring number -> [(atom, bond symbol)] in order of occurrence
    start = True
    i, n = 0, len(s)

    def add_atom(atom: Atom):
        nonlocal prev, pending, start
        mol._atoms.append(atom)
        atom._smiles_start = start
        start = False
        if prev is not None:
            if pending in ("/", "\\"):
                order, direction = 1.0, pending
            elif pending:
                order, direction = _ORDER[pending], ""
            else:
                order, direction = (1.5 if (prev._aromatic and atom._aromatic) else 1.0), ""
            mol._add_bond(prev, atom, order, direction)
        elif pending:
            raise SmilesError("bond without a preceding atom")
        prev, pending = atom, ""

    def ring(num: int):
        nonlocal pending
        if prev is None:
            raise SmilesError("ring closure without an atom")
        open_rings.setdefault(num, []).append((prev, pending))
        prev._ring_closures.append((num, len(open_rings[num]) - 1))
        pending = ""

    while i < n:
        ch = s[i]
        if ch == "[":
            j = s.find("]", i)
            if j < 0:
                raise SmilesError("unclosed [")
            add_atom(_parse_bracket(mol, s[i + 1:j]))
            i = j + 1
        elif s[i:i + 2] in ("Cl", "Br"):
            add_atom(Atom(mol, len(mol._atoms), s[i:i + 2])); i += 2
        elif ch in _ORGANIC:
            add_atom(Atom(mol, len(mol._atoms), ch)); i += 1
        elif ch in _AROMATIC_ORGANIC:
            add_atom(Atom(mol, len(mol._atoms), _AROMATIC_ORGANIC[ch], aromatic=True)); i += 1
        elif ch in _ORDER or ch in "/\\":
            if pending:
                raise SmilesError("two bond symbols in a row")
            pending = ch; i += 1
        elif ch == "(":
            if prev is None:
                raise SmilesError("branch without an atom")
            stack.append(prev); i += 1
        elif ch == ")":
            if not stack or pending:
                raise SmilesError("unbalanced )")
            prev = stack.pop(); i += 1
        elif ch.isdigit():
            ring(int(ch)); i += 1
        elif ch == "%":
            if not s[i + 1:i + 3].isdigit():
                raise SmilesError("bad %nn ring closure")
            ring(int(s[i + 1:i + 3])); i += 3
        elif ch == ".":
            if pending:
                raise SmilesError("dangling bond before .")
            prev, start = None, True; i += 1
        else:
            raise SmilesError(f"unexpected character {ch!r}")
    if stack or pending:
        raise SmilesError("unbalanced ( or dangling bond")
    # ring-closure bonds: after everything else, by ring number, pairs in order of occurrence
    made: Dict[Tuple[int, int], Bond] = {}
    for num in sorted(open_rings):
        uses = open_rings[num]
        if len(uses) % 2:
            raise SmilesError(f"unclosed ring {num}")
        for k in range(0, len(uses), 2):
            (a, sa), (b, sb) = uses[k], uses[k + 1]
            if a is b or mol.GetBondBetweenAtoms(a._idx, b._idx) is not None:
                raise SmilesError(f"bad ring closure {num}")
            direction = ""
            if sa in ("/", "\\"):
                direction, sa = sa, ""
            elif sb in ("/", "\\"):        # written at the closing atom: seen from the other end
                direction, sb = ("\\" if sb == "/" else "/"), ""
            if sa and sb and _ORDER[sa] != _ORDER[sb]:
                raise SmilesError(f"ring closure {num}: bond orders disagree")
            sym = sa or sb
            order = _ORDER[sym] if sym else (1.5 if (a._aromatic and b._aromatic) else 1.0)
            bond = mol._add_bond(a, b, order, direction)
            made[(num, k)] = bond
            made[(num, k + 1)] = bond
    for atom in mol._atoms:
        atom._ring_closures = [made[key] for key in atom._ring_closures]
    if not mol._atoms:
        raise SmilesError("no atoms")
    if len(mol._atoms) > 1 and any(a._symbol == "H" for a in mol._atoms):
        # RDKit folds [H] atoms into their neighbours (removeHs, with its own rules for isotopes, chirality and
        # bridging hydrogens); not restated here, so refuse instead of numbering a different graph
        raise SmilesError("explicit hydrogen atoms need RDKit")
    return mol


# ------------------------------------------------------------- perception ----

def _kekulize(mol: Mol) -> None:
    """Aromatic (lower-case) input -> one Kekule structure; the flags are re-perceived afterwards."""
    arom_bonds = [b for b in mol._bonds if b._order == 1.5]
    if not arom_bonds:
        for a in mol._atoms:
            a._aromatic = False
        return
    need = set()
    for a in mol._atoms:
        if not a._aromatic:
            continue
        sigma = len(a._bonds) + a._explicit_hs
        if a._symbol in ("C", "B"):
            wants = a._charge == 0
        elif a._symbol in ("N", "P"):
            wants = (sigma == 2 and a._charge == 0) or (sigma == 3 and a._charge == 1)
        elif a._symbol in ("O", "S", "Se"):
            wants = sigma == 2 and a._charge == 1          # pyrylium / thiopyrylium
        else:
            wants = False
        if wants and not any(b._order >= 2.0 for b in a._bonds):
            need.add(a._idx)
    adj = {i: [] for i in need}
    for b in arom_bonds:
        if b._a._idx in need and b._b._idx in need:
            adj[b._a._idx].append((b._b._idx, b))
            adj[b._b._idx].append((b._a._idx, b))
    order = sorted(need)
    matched: Dict[int, Bond] = {}

    def solve(k: int) -> bool:
        while k < len(order) and order[k] in matched:
            k += 1
        if k == len(order):
            return True
        u = order[k]
        for v, b in adj[u]:
            if v not in matched:
                matched[u] = matched[v] = b
                if solve(k + 1):
                    return True
                del matched[u], matched[v]
        return False

    if not solve(0):
        raise SmilesError("cannot kekulize")
    doubles = {b._idx for b in matched.values()}
    for b in arom_bonds:
        b._order = 2.0 if b._idx in doubles else 1.0
    for a in mol._atoms:
        a._aromatic = False


def _valences(mol: Mol) -> None:
    """Explicit valence, implicit hydrogens, radicals (RDKit Atom::calcExplicitValence / calcImplicitValence
    on the Kekule form; MolOps::assignRadicals for bracket atoms)."""
    for a in mol._atoms:
        accum = sum(b._order for b in a._bonds) + a._explicit_hs
        a._explicit_valence = int(accum + 0.1 + 0.5)
        if a._no_implicit:
            a._implicit_hs = 0
            base = 2 if a._z <= 2 else 8
            rad = base - a._nouter - a._explicit_valence + a._charge
            if rad < 0:
                rad = 0
                if len(a._valences) > 1:
                    for v in a._valences:
                        if v - a._explicit_valence + a._charge >= 0:
                            rad = v - a._explicit_valence + a._charge
                            break
            rad2 = a._nouter - a._explicit_valence - a._charge
            if rad2 >= 0:
                rad = min(rad, rad2)
            a._radicals = rad
            continue
        chg = a._charge
        if a._z == 6 and chg > 0:
            chg = -chg
        if a._z in (5,):                    # early atoms: the charge correction is reversed
            chg = -chg
        a._implicit_hs = 0
        for v in a._valences:
            if a._explicit_valence <= v + chg:
                a._implicit_hs = v + chg - a._explicit_valence
                break
        else:
            raise SmilesError(f"valence of atom {a._idx} ({a._symbol}) exceeds the allowed maximum")


def _find_rings(mol: Mol) -> None:
    """Ring bonds (non-bridges) and, for every ring bond, all shortest cycles through it: the ring set used
    for ring sizes and aromaticity (the symmetrised smallest set of smallest rings for the ring systems of
    interest here: fused polycycles, bicyclics, small cages)."""
    n = len(mol._atoms)
    nbrs = [[(b._other(a)._idx, b._idx) for b in a._bonds] for a in mol._atoms]
    rings: Dict[frozenset, Tuple[int, ...]] = {}
    for bond in mol._bonds:
        s, t = bond._a._idx, bond._b._idx
        # BFS from s avoiding the bond itself; collect every shortest path to t
        dist = [-1] * n
        preds: List[List[Tuple[int, int]]] = [[] for _ in range(n)]
        dist[s] = 0
        frontier = [s]
        while frontier and dist[t] < 0:
            nxt = []
            for u in frontier:
                for v, bi in nbrs[u]:
                    if bi == bond._idx:
                        continue
                    if dist[v] < 0:
                        dist[v] = dist[u] + 1
                        nxt.append(v)
                    if dist[v] == dist[u] + 1:
                        preds[v].append((u, bi))
            frontier = nxt
        if dist[t] < 0:
            continue
        bond._in_ring = True
        if dist[t] + 1 > 24 and len(rings) > 64:
            continue
        paths: List[Tuple[List[int], List[int]]] = []

        def walk(v: int, atoms: List[int], bonds: List[int]):
            if len(paths) > 64:
                return
            if v == s:
                paths.append((atoms[:], bonds[:]))
                return
            for u, bi in preds[v]:
                atoms.append(u); bonds.append(bi)
                walk(u, atoms, bonds)
                atoms.pop(); bonds.pop()

        walk(t, [t], [bond._idx])
        for atoms, bonds in paths:
            key = frozenset(bonds)
            if key not in rings:
                rings[key] = tuple(atoms)
    items = sorted(rings.items(), key=lambda kv: (len(kv[1]), sorted(kv[1])))
    mol._bond_rings = [k for k, _ in items]
    mol._rings = [v for _, v in items]
    for a in mol._atoms:
        sizes = [len(r) for r in mol._rings if a._idx in r]
        a._ring_size = min(sizes) if sizes else 0


_VACANT, _ONE, _TWO, _NONE = 0, 1, 2, -1


def _count_atom_elec(a: Atom) -> int:
    """RDKit countAtomElec: electrons an atom can give to a pi system."""
    dv = a._valences[0]
    if dv <= 1:
        return 0
    degree = len(a._bonds) + a.GetTotalNumHs()
    if degree > 3:
        return 0
    nlp = max((a._nouter - dv) - a._charge, 0)
    res = (dv - degree) + nlp - a._radicals
    if res > 1 and a._explicit_valence - len(a._bonds) > 1:
        res = 1
    return res


def _more_electronegative(x: Atom, y: Atom) -> bool:
    return x._nouter > y._nouter or (x._nouter == y._nouter and x._z < y._z)


def _donor_type(a: Atom) -> int:
    nelec = _count_atom_elec(a)
    exo = next((b for b in a._bonds if not b._in_ring and b._order >= 2.0), None)
    if nelec < 0:
        return _NONE
    if nelec == 0:
        if exo is not None:
            return _VACANT
        if any(b._in_ring and b._order >= 2.0 for b in a._bonds):
            return _ONE
        return _NONE
    if nelec == 1:
        if exo is not None:
            return _VACANT if _more_electronegative(exo._other(a), a) else _ONE
        if a._explicit_valence != len(a._bonds) + a._explicit_hs:
            return _ONE
        if a._charge == 1:
            return _VACANT
        return _NONE
    if exo is not None and _more_electronegative(exo._other(a), a):
        nelec -= 1
    return _ONE if nelec % 2 == 1 else _TWO


def _arom_candidate(a: Atom, donor: int) -> bool:
    if a._z > 18 and a._z not in (34, 52):
        return False
    if donor == _NONE:
        return False
    iso = _BY_NUMBER.get(a._z - a._charge)
    if iso is not None and a.GetTotalValence() > iso[2][0]:
        return False
    if a._radicals and (a._z != 6 or a._charge):
        return False
    if a._explicit_valence - len(a._bonds) > 1 and sum(1 for b in a._bonds if b._order >= 2.0) > 1:
        return False
    return True


def _huckel(atoms, donor) -> bool:
    total = sum(donor[i] for i in atoms)
    if total >= 6:
        return (total - 2) % 4 == 0
    return total == 2


def _aromaticity(mol: Mol) -> None:
    """RDKit's default model on the Kekule form: candidate rings, then Hueckel on single rings and on fused
    combinations (up to six rings); in a combination only atoms in one or two of its rings count, and only
    bonds in exactly one of its rings are marked."""
    donor = {a._idx: _donor_type(a) for a in mol._atoms if a._ring_size}
    cand = {i: _arom_candidate(mol._atoms[i], d) for i, d in donor.items()}
    ring_ids = [k for k, r in enumerate(mol._rings) if all(cand[i] for i in r)]
    if not ring_ids:
        return
    fused = {k: set() for k in ring_ids}
    for x, y in itertools.combinations(ring_ids, 2):
        if len(mol._rings[x]) <= 24 and len(mol._rings[y]) <= 24 and mol._bond_rings[x] & mol._bond_rings[y]:
            fused[x].add(y); fused[y].add(x)
    seen = set()
    for root in ring_ids:
        if root in seen:
            continue
        system, todo = [], [root]
        seen.add(root)
        while todo:
            k = todo.pop()
            system.append(k)
            for m in fused[k]:
                if m not in seen:
                    seen.add(m); todo.append(m)
        system.sort()
        all_bonds = set().union(*(mol._bond_rings[k] for k in system))
        done = set()
        for size in range(1, min(len(system), 6) + 1):
            if done >= all_bonds:
                break
            if size > 2 and len(system) > 300:
                break
            for combo in itertools.combinations(system, size):
                if size > 1:                # the chosen rings must hang together
                    reach, todo = {combo[0]}, [combo[0]]
                    while todo:
                        k = todo.pop()
                        for m in fused[k]:
                            if m in combo and m not in reach:
                                reach.add(m); todo.append(m)
                    if len(reach) != size:
                        continue
                count: Dict[int, int] = {}
                for k in combo:
                    for i in mol._rings[k]:
                        count[i] = count.get(i, 0) + 1
                if not _huckel([i for i, c in count.items() if c in (1, 2)], donor):
                    continue
                bcount: Dict[int, int] = {}
                for k in combo:
                    for i in mol._rings[k]:
                        mol._atoms[i]._aromatic = True
                    for bi in mol._bond_rings[k]:
                        bcount[bi] = bcount.get(bi, 0) + 1
                for bi, c in bcount.items():
                    if c == 1:
                        mol._bonds[bi]._aromatic = True
                        done.add(bi)
    for b in mol._bonds:
        if b._aromatic and b._order in (1.0, 2.0):
            b._order = 1.5


def _conjugation(mol: Mol) -> None:
    """RDKit MolOps::setConjugation / markConjAtomBonds."""
    for b in mol._bonds:
        b._conjugated = b._aromatic
    for at in mol._atoms:
        sbo = len(at._bonds) + at.GetTotalNumHs()
        if sbo < 2 or sbo > 3:
            continue
        for b1 in at._bonds:
            if b1._order < 1.5:
                continue
            for b2 in at._bonds:
                if b2 is b1:
                    continue
                at2 = b2._other(at)
                if len(at2._bonds) + at2.GetTotalNumHs() > 3:
                    continue
                if _count_atom_elec(at2) > 0:
                    b1._conjugated = True
                    b2._conjugated = True


def _hybridization(mol: Mol) -> None:
    """RDKit MolOps::setHybridization (numBondsPlusLonePairs)."""
    for a in mol._atoms:
        deg = a.GetTotalDegree()
        tv = a.GetTotalValence()
        free = a._nouter - (tv + a._charge)
        if tv + a._nouter - a._charge < 8:
            norbs = deg + (free - a._radicals) // 2 + a._radicals
        else:
            norbs = deg + free // 2
        if norbs <= 1:
            a._hyb = HybridizationType.S
        elif norbs == 2:
            a._hyb = HybridizationType.SP
        elif norbs == 3:
            a._hyb = HybridizationType.SP2
        elif norbs == 4:
            conj = any(b._conjugated for b in a._bonds)
            a._hyb = HybridizationType.SP3 if (len(a._bonds) > 3 or not conj) else HybridizationType.SP2
        elif norbs == 5:
            a._hyb = HybridizationType.SP3D
        elif norbs == 6:
            a._hyb = HybridizationType.SP3D2
        else:
            a._hyb = HybridizationType.UNSPECIFIED


def _cip_ranks(mol: Mol) -> List[int]:
    """Iterative CIP-style ranks (RDKit Chirality::assignAtomCIPRanks): start from (atomic number, isotope);
    each round appends the neighbours' ranks, highest first, every neighbour 2 x bond order times, a zero per
    hydrogen; stop when the number of classes no longer grows."""
    n = len(mol._atoms)

    def dense(keys):
        order = {k: r for r, k in enumerate(sorted(set(keys)))}
        return [order[k] for k in keys]

    ranks = dense([(a._z, a._isotope) for a in mol._atoms])
    n_classes = len(set(ranks))
    for _ in range(max(n // 2, 1) + 1):
        if n_classes == n:
            break
        keys = []
        for a in mol._atoms:
            entry = []
            for b in a._bonds:
                nbr = b._other(a)
                if b._order == 2.0 and nbr._z == 15 and len(nbr._bonds) in (3, 4):
                    reps = 1
                else:
                    reps = int(2.0 * b._order + 0.1)
                entry.extend([ranks[nbr._idx] + 1] * reps)
            entry.extend([0] * a.GetTotalNumHs())
            entry.sort(reverse=True)
            keys.append((ranks[a._idx], tuple(entry)))
        new = dense(keys)
        k = len(set(new))
        ranks = new
        if k == n_classes:
            break
        n_classes = k
    return ranks


def _permutation_parity(seq: List[int], ref: List[int]) -> int:
    """Number of swaps (mod 2) that turn `ref` into `seq` (both hold the same distinct items)."""
    pos = {v: i for i, v in enumerate(ref)}
    perm = [pos[v] for v in seq]
    swaps, seen = 0, [False] * len(perm)
    for i in range(len(perm)):
        if seen[i]:
            continue
        j, length = i, 0
        while not seen[j]:
            seen[j] = True
            j = perm[j]
            length += 1
        swaps += length - 1
    return swaps % 2


def _stereo(mol: Mol) -> None:
    ranks = None
    # ---- tetrahedral centres: '@' = counter-clockwise in SMILES order; RDKit stores the tag relative to the
    # atom's bond order, in which ring closures come last (SmilesParseOps::AdjustAtomChiralityFlags)
    tagged = [a for a in mol._atoms if a._chiral_mark]
    for a in tagged:
        ccw = a._chiral_mark == "@"
        closures = [b._idx for b in a._ring_closures]
        chain = sorted((b._other(a)._idx, b._idx) for b in a._bonds if b._idx not in closures)
        smiles_order = [bi for j, bi in chain if j < a._idx] + closures + [bi for j, bi in chain if j > a._idx]
        swaps = _permutation_parity(smiles_order, [b._idx for b in a._bonds])
        if len(a._bonds) == 3 and a._smiles_start and a._explicit_hs == 1:
            swaps += 1
        if swaps % 2:
            ccw = not ccw
        a._tag = ChiralType.CHI_TETRAHEDRAL_CCW if ccw else ChiralType.CHI_TETRAHEDRAL_CW
    if tagged:
        ranks = _cip_ranks(mol)
        keep = set()
        maybe_ring = []
        for a in tagged:
            total = a.GetTotalDegree()
            legal = 3 <= total <= 4 and len(a._bonds) >= 3 and a.GetTotalNumHs() <= 1
            if a._z == 7 and total == 3 and a._ring_size != 3:
                legal = False                  # a three-coordinate nitrogen inverts (RDKit keeps it only in 3-rings / bridgeheads)
            nbr_ranks = [ranks[x._idx] for x in a._neighbors()]
            dupes = len(set(nbr_ranks)) != len(nbr_ranks)
            if legal and not dupes:
                keep.add(a._idx)
            elif legal and dupes and a._ring_size:
                ring_nbrs = [b for b in a._bonds if b._in_ring]
                other = [b._other(a) for b in a._bonds if not b._in_ring]
                if (len(other) == 1 and len(ring_nbrs) >= 2) or (len(other) == 0 and len(ring_nbrs) == 4) or \
                        (len(other) == 2 and ranks[other[0]._idx] != ranks[other[1]._idx]):
                    maybe_ring.append(a)
        ring_tagged = keep | {a._idx for a in maybe_ring}
        for a in maybe_ring:                # ring stereo needs a partner in one of its rings
            if any(a._idx in r and any(j in ring_tagged and j != a._idx for j in r) for r in mol._rings):
                keep.add(a._idx)
        for a in tagged:
            if a._idx not in keep:
                a._tag = ChiralType.CHI_UNSPECIFIED
    # ---- double bonds
    if not any(b._dir for b in mol._bonds):
        return
    if ranks is None:
        ranks = _cip_ranks(mol)
    for bond in mol._bonds:
        if bond._order != 2.0 or bond._aromatic:
            continue
        if bond._in_ring and min(len(r) for r in mol._bond_rings if bond._idx in r) < 8:
            continue
        sides = []
        for end, far in ((bond._a, bond._b), (bond._b, bond._a)):
            subs = [b for b in end._bonds if b is not bond]
            if not 1 <= len(subs) <= 2:
                break
            if len(subs) == 2 and ranks[subs[0]._other(end)._idx] == ranks[subs[1]._other(end)._idx]:
                break
            up: Dict[int, bool] = {}
            for b in subs:
                if b._order != 1.0 or not b._dir:
                    continue
                rising = b._dir == "/"      # the end atom of the bond lies above its begin atom
                up[b._idx] = rising if b._a is end else not rising
            if not up:
                break
            if len(up) == 2 and len(set(up.values())) != 2:
                break                       # both substituents on the same side: contradictory marks
            for b in subs:
                if b._idx not in up:
                    up[b._idx] = not next(iter(up.values()))
            top = max(subs, key=lambda b: ranks[b._other(end)._idx])
            sides.append(up[top._idx])
        if len(sides) == 2:
            bond._stereo = BondStereo.STEREOZ if sides[0] == sides[1] else BondStereo.STEREOE


def _perceive(mol: Mol) -> Mol:
    _kekulize(mol)
    _valences(mol)
    _find_rings(mol)
    _aromaticity(mol)
    _conjugation(mol)
    _hybridization(mol)
    _stereo(mol)
    return mol


def MolFromSmiles(smiles: str) -> Optional[Mol]:
    """RDKit's contract: a perceived molecule, or None when the string cannot be parsed / sanitised."""
    try:
        return _perceive(_parse(smiles))
    except (SmilesError, KeyError, IndexError):
        return None
