"""-m "not gpu": the RDKit-free featuriser (graphchem_b200/preprocessing/smiles.py).

Pins, all integers / strings produced by the real reference stack and committed here:
* the reference's own test strings and shapes (reference tests/test_features.py:14-168, replayed below with
  `Chem` = the built-in module);
* the four notebook encodings (tests/golden/encodings.json);
* the eight vocabulary sizes the notebooks print for `MoleculeEncoder(smiles_train)` on the CN / YSI / RON / LHV
  training splits (tests/golden/fuel_sets.json) -- first-seen ids depend on the PARTITION of all atoms / bonds of
  ~1400 molecules into property classes, so equal sizes plus equal first-molecule ids pin the perception rules
  (valence, rings, aromaticity, conjugation, hybridisation, stereo) on every chemistry the data sets contain."""
import json
import os

import pytest
import torch

from graphchem_b200.preprocessing import MoleculeEncoder, atom_to_str, bond_to_str, get_ring_size, load_encoder
from graphchem_b200.preprocessing import smiles as Chem
from tests.util import load_fuel_set

HERE = os.path.dirname(os.path.abspath(__file__))
BENZENE_ATOM = ("(rdkit.Chem.rdchem.ChiralType.CHI_UNSPECIFIED, 2, 3, 0, rdkit.Chem.rdchem.HybridizationType.SP2, 1, "
                "True, False, 0, 1, 0, 'C', 3, 1, 4, 6)")
BENZENE_BOND = ("(rdkit.Chem.rdchem.BondType.AROMATIC, True, rdkit.Chem.rdchem.BondStereo.STEREONONE, 6, "
                "[['C', 'C']])")


# ---- the reference's tests/test_features.py, replayed ----

def test_reference_get_ring_size():
    mol = Chem.MolFromSmiles("C1=CC=CC=C1")
    assert all(get_ring_size(a) == 6 for a in mol.GetAtoms())
    assert all(get_ring_size(a) == 0 for a in Chem.MolFromSmiles("CC").GetAtoms())
    for a in mol.GetAtoms():
        assert (get_ring_size(a, max_size=5), get_ring_size(a, max_size=6), get_ring_size(a, max_size=7)) == (5, 6, 6)


def test_reference_atom_and_bond_strings_for_benzene():
    mol = Chem.MolFromSmiles("C1=CC=CC=C1")
    assert mol.GetNumAtoms() == 6 and mol.GetNumBonds() == 6
    for i in range(6):
        assert atom_to_str(mol.GetAtomWithIdx(i)) == BENZENE_ATOM
        assert bond_to_str(mol.GetBondWithIdx(i)) == BENZENE_BOND


def test_reference_encoder_shapes_and_persistence(tmp_path):
    enc = MoleculeEncoder(["C", "N"], backend="builtin")
    assert len(enc.vocab_sizes) == 2 and min(enc.vocab_sizes) > 1
    for atoms, bonds, conn in enc.encode_many(["C", "N"]):
        assert len(atoms) > 0 and conn.shape == (2, len(bonds))
    enc2 = MoleculeEncoder(["C=C"], backend="builtin")
    a, b, c = enc2.encode("C=C")
    assert a.shape == (2,) and b.shape == (2,) and c.shape == (2, 2)
    assert (a.dtype, b.dtype, c.dtype) == (torch.int32, torch.int32, torch.int64)
    enc.save(str(tmp_path / "e.pkl"))
    loaded = load_encoder(str(tmp_path / "e.pkl"))
    assert loaded.vocab_sizes == enc.vocab_sizes
    assert torch.equal(loaded.encode("CN")[0], enc.encode("CN")[0])
    with pytest.raises(ValueError):
        MoleculeEncoder(["C", "not a smiles"], backend="builtin")
    with pytest.raises(ValueError):
        enc.encode("C1CC")


# ---- notebook pins ----

@pytest.mark.parametrize("name", ["cn", "ysi", "ron", "mon", "lhv"])
def test_notebook_vocabulary_sizes_and_first_encoding(name):
    smiles, y, tr, te, vocab = load_fuel_set(name)
    train = [smiles[i] for i in tr]
    enc = MoleculeEncoder(train, backend="builtin")
    assert enc.vocab_sizes == vocab
    gold = json.load(open(os.path.join(HERE, "golden", "encodings.json")))
    key = {"cn": "cn", "ysi": "ysi", "ron": "ron", "mon": "ron", "lhv": "lhv"}[name]
    entry = next(v for k, v in gold.items() if k.startswith(key + ":") and k.split(":", 1)[1] == train[0])
    atoms, bonds, conn = enc.encode(train[0])
    assert atoms.tolist() == entry["atoms"] and bonds.tolist() == entry["bonds"]
    assert conn.tolist() == entry["connectivity"]
    # every molecule of the set encodes; the test split meets no unknown atom class the notebooks would have hidden
    for s in smiles:
        a, b, c = enc.encode(s)
        assert a.min() >= 1 and (b.numel() == 0 or b.min() >= 1)
        assert a.max() < vocab[0] and (b.numel() == 0 or b.max() < vocab[1])
        assert c.shape == (2, b.numel()) and b.numel() % 2 == 0
        src, dst = c[0].tolist(), c[1].tolist()
        assert src == sorted(src)                                     # source-atom-major (features.py:302-317)
        rev = {(s_, d_): t for s_, d_, t in zip(src, dst, b.tolist())}
        assert all(rev[(d_, s_)] == t for (s_, d_), t in rev.items())   # both directions share the bond token


# ---- perception rules the pins exercise, stated directly ----

def _strings(smi):
    mol = Chem.MolFromSmiles(smi)
    assert mol is not None, smi
    return sorted(atom_to_str(a) for a in mol.GetAtoms()), sorted(bond_to_str(b) for b in mol.GetBonds())


@pytest.mark.parametrize("kekule,aromatic", [
    ("C1=CC=CC=C1", "c1ccccc1"), ("CC(C)(C)C1=CC=CO1", "CC(C)(C)c1ccco1"), ("C1=CC=NC=C1", "c1ccncc1"),
    ("C1=CC=C2C=CC=CC2=C1", "c1ccc2ccccc2c1"), ("C1=CSC=C1", "c1ccsc1"), ("CN1C=CC2=CC=CC=C21", "Cn1ccc2ccccc21"),
])
def test_kekule_and_aromatic_spellings_agree(kekule, aromatic):
    assert _strings(kekule) == _strings(aromatic)
    assert all(a.GetIsAromatic() for a in Chem.MolFromSmiles(aromatic).GetAtoms() if a.IsInRing())


def test_aromaticity_model():
    arom = lambda s: [a.GetIsAromatic() for a in Chem.MolFromSmiles(s).GetAtoms()]  # noqa: E731
    assert not any(arom("C1=CC=CC=CC=C1"))            # cyclooctatetraene: 8 electrons
    assert not any(arom("C1=CCC=C1"))                 # cyclopentadiene: sp3 carbon in the ring
    assert not any(arom("O=C1C=CC(=O)C=C1"))          # quinone: exocyclic C=O steals the electron
    assert all(arom("O=C1C=CC=CO1")[1:])              # 2-pyranone is aromatic in RDKit's model
    assert all(arom("C1=CC2=C3C(=C1)C=CC4=CC=CC(=C43)C=C2"))   # pyrene
    fl = Chem.MolFromSmiles("C1=CC=C2C(=C1)C3=CC=CC4=C3C2=CC=C4")  # fluoranthene: the five-ring's own bonds stay single
    assert all(a.GetIsAromatic() for a in fl.GetAtoms())
    assert sorted(str(b.GetBondType()) for b in fl.GetBonds()).count("SINGLE") == 2
    ind = Chem.MolFromSmiles("C1C=CC2=CC=CC=C12")     # indene: only the benzene ring
    assert [a.GetIsAromatic() for a in ind.GetAtoms()] == [False, False, False, True, True, True, True, True, True]


def test_conjugation_and_hybridisation():
    mol = Chem.MolFromSmiles("CCOC(=O)C")             # ester: C-O and C=O conjugated, ether O sp2, alkyl C-C not
    conj = {(b.GetBeginAtomIdx(), b.GetEndAtomIdx()): b.GetIsConjugated() for b in mol.GetBonds()}
    assert conj == {(0, 1): False, (1, 2): False, (2, 3): True, (3, 4): True, (3, 5): False}
    assert [str(a.GetHybridization()) for a in mol.GetAtoms()] == ["SP3", "SP3", "SP2", "SP2", "SP2", "SP3"]
    assert [str(a.GetHybridization()) for a in Chem.MolFromSmiles("CC#N").GetAtoms()] == ["SP3", "SP", "SP"]
    assert str(Chem.MolFromSmiles("NC1=CC=CC=C1").GetAtomWithIdx(0).GetHybridization()) == "SP2"   # aniline N
    assert str(Chem.MolFromSmiles("NCC").GetAtomWithIdx(0).GetHybridization()) == "SP3"
    nitro = Chem.MolFromSmiles("C[N+](=O)[O-]")
    assert [a.GetFormalCharge() for a in nitro.GetAtoms()] == [0, 1, 0, -1]
    assert [a.GetTotalNumHs() for a in nitro.GetAtoms()] == [3, 0, 0, 0]
    assert [a.GetNoImplicit() for a in nitro.GetAtoms()] == [False, True, False, True]
    co = Chem.MolFromSmiles("[C-]#[O+]")
    assert [str(a.GetHybridization()) for a in co.GetAtoms()] == ["SP", "SP"]


def test_bond_order_of_ring_closures():
    """Chain bonds in reading order, ring-closure bonds afterwards by ring number (RDKit CloseMolRings):
    this is what orders a molecule's edges (App. B: 'ring-closure bonds come after chain bonds')."""
    mol = Chem.MolFromSmiles("C1CCCCC1C")
    assert [(b.GetBeginAtomIdx(), b.GetEndAtomIdx()) for b in mol.GetBonds()] == \
        [(0, 1), (1, 2), (2, 3), (3, 4), (4, 5), (5, 6), (0, 5)]
    mol = Chem.MolFromSmiles("C2CC1CCC1C2")
    assert [(b.GetBeginAtomIdx(), b.GetEndAtomIdx()) for b in mol.GetBonds()][-2:] == [(2, 5), (0, 6)]
    assert [b.GetIdx() for b in mol.GetAtomWithIdx(5).GetBonds()] == [4, 5, 6]
    mol = Chem.MolFromSmiles("C%12CC%12")
    assert mol.GetNumBonds() == 3 and all(get_ring_size(a) == 3 for a in mol.GetAtoms())
    cage = Chem.MolFromSmiles("C1C2C3C2C4C1C34")      # lhv.csv: ring sizes through fused three-rings
    assert sorted(get_ring_size(a) for a in cage.GetAtoms()) == [3, 3, 3, 3, 3, 3, 5]


def test_double_bond_stereo():
    st = lambda s: [str(b.GetStereo()) for b in Chem.MolFromSmiles(s).GetBonds() if str(b.GetBondType()) == "DOUBLE"]  # noqa: E731
    assert st("F/C=C/F") == ["STEREOE"] and st("F/C=C\\F") == ["STEREOZ"] and st("F\\C=C\\F") == ["STEREOE"]
    assert st("C(/F)=C/F") == ["STEREOZ"] and st("FC=CF") == ["STEREONONE"]
    assert st("C/C=C(/C)C") == ["STEREONONE"]                          # two equal substituents: no stereo
    assert st("CC/C(=C/C)/C") == ["STEREOE"] or st("CC/C(=C/C)/C") == ["STEREOZ"]
    assert st("CC/C(=C/C)/C") != st("CC/C(=C\\C)/C")
    assert st("C1CCC/C=C\\CC1") == ["STEREOZ"]                         # eight-ring: stereo kept
    assert st("C1CC/C=C\\C1") == ["STEREONONE"]                        # rings below 8 carry no double-bond stereo
    assert st("C/C=C/C=C/C") == ["STEREOE", "STEREOE"] and st("C/C=C\\C=C/C") == ["STEREOZ", "STEREOZ"] and st("C/C=C/C=C\\C") == ["STEREOE", "STEREOZ"]
    assert st("CC\\C=C/C\\C=C/CC") == ["STEREOZ", "STEREOZ"]           # cn.csv fatty esters


def test_tetrahedral_tags():
    tag = lambda s, i: str(Chem.MolFromSmiles(s).GetAtomWithIdx(i).GetChiralTag())  # noqa: E731
    assert {tag("N[C@@H](C)O", 1), tag("N[C@H](C)O", 1)} == {"CHI_TETRAHEDRAL_CW", "CHI_TETRAHEDRAL_CCW"}
    assert tag("N[C@@H](C)O", 1) == "CHI_TETRAHEDRAL_CW" and tag("N[C@](F)(C)O", 1) == "CHI_TETRAHEDRAL_CCW"
    # the same enantiomer written with the hydrogen-bearing atom first: the implicit H becomes the 'from' atom
    assert tag("[C@@H](N)(C)O", 0) != tag("N[C@@H](C)O", 1)
    # a ring closure sits before the chain neighbours in the SMILES but after them in the bond list
    assert tag("C[C@H]1CCCO1", 1) != tag("C[C@H](CCC1)O1", 1)
    # not a stereocentre (two methyls): the tag is dropped, the written hydrogen stays explicit
    mol = Chem.MolFromSmiles("C[C@H](C)O")
    a = mol.GetAtomWithIdx(1)
    assert str(a.GetChiralTag()) == "CHI_UNSPECIFIED" and a.GetNumExplicitHs() == 1 and a.GetNoImplicit()
    assert atom_to_str(a) != atom_to_str(Chem.MolFromSmiles("CC(C)O").GetAtomWithIdx(1))
    # ring stereo: 1,4-dimethylcyclohexane has no true stereocentre but keeps both tags (cis / trans)
    mol = Chem.MolFromSmiles("C[C@H]1CC[C@@H](C)CC1")
    assert [str(a.GetChiralTag()) for a in mol.GetAtoms()].count("CHI_UNSPECIFIED") == 6


@pytest.mark.parametrize("bad", ["", "C(", "C)", "C1CC", "C=", "C(C)(C)(C)(C)C", "[C", "c1cccc1", "C%1CC", "Xx"])
def test_unparsable_smiles_give_none(bad):
    assert Chem.MolFromSmiles(bad) is None


def test_azulene_keeps_its_fusion_bond_single():
    """RDKit writes azulene as c1ccc2cccc-2cc1: neither ring is aromatic alone (5 and 7 electrons), the 10-electron
    envelope is, and only envelope bonds are marked."""
    mol = Chem.MolFromSmiles("C1=CC2=CC=CC=CC2=C1")
    assert all(a.GetIsAromatic() for a in mol.GetAtoms())
    kinds = [str(b.GetBondType()) for b in mol.GetBonds()]
    assert kinds.count("AROMATIC") == 10 and kinds.count("SINGLE") == 1
    assert _strings("C1=CC2=CC=CC=CC2=C1") == _strings("c1ccc2cccc-2cc1")


def _write_random_smiles(mol, rng):
    """A SMILES of `mol` from a random start atom with random branch order, aromatic atoms in lower case (no stereo)."""
    n = mol.GetNumAtoms()
    nbrs = {a.GetIdx(): [(b.GetEndAtomIdx() if b.GetBeginAtomIdx() == a.GetIdx() else b.GetBeginAtomIdx(), b)
                         for b in a.GetBonds()] for a in mol.GetAtoms()}
    for v in nbrs.values():
        rng.shuffle(v)
    start = int(rng.integers(n))
    # pass 1: DFS tree + ring-closure bonds
    seen, tree, order = {start}, set(), []
    closures = {i: [] for i in range(n)}
    stack = [(start, None)]
    parent_bond = {start: None}

    def visit(u):
        order.append(u)
        for v, b in nbrs[u]:
            if b.GetIdx() == (parent_bond[u].GetIdx() if parent_bond[u] is not None else -1):
                continue
            if v in seen:
                if b.GetIdx() not in tree and all(b.GetIdx() != c.GetIdx() for c in closures[u]):
                    closures[u].append(b)
                    closures[v].append(b)
            else:
                seen.add(v); tree.add(b.GetIdx()); parent_bond[v] = b
                visit(v)

    visit(start)
    assert len(seen) == n
    digit = {}

    def bond_symbol(b):
        a1, a2 = b.GetBeginAtom(), b.GetEndAtom()
        kind = str(b.GetBondType())
        if kind == "AROMATIC":
            return ""
        if kind == "SINGLE":
            return "-" if (a1.GetIsAromatic() and a2.GetIsAromatic()) else ""
        return {"DOUBLE": "=", "TRIPLE": "#"}[kind]

    def atom_symbol(a):
        sym = a.GetSymbol().lower() if a.GetIsAromatic() else a.GetSymbol()
        if a.GetNoImplicit() or a.GetFormalCharge():
            h = a.GetNumExplicitHs()
            c = a.GetFormalCharge()
            return "[" + sym + ("H" + (str(h) if h > 1 else "") if h else "") + ("+" * c if c > 0 else "-" * -c) + "]"
        return sym

    out = []

    def emit(u):
        out.append(atom_symbol(mol.GetAtomWithIdx(u)))
        for b in closures[u]:
            if b.GetIdx() not in digit:
                digit[b.GetIdx()] = len(digit) + 1
                out.append(bond_symbol(b))
            k = digit[b.GetIdx()]
            out.append(str(k) if k < 10 else f"%{k}")
        kids = [(v, b) for v, b in nbrs[u] if parent_bond.get(v) is b and v != start]
        for i, (v, b) in enumerate(kids):
            last = i == len(kids) - 1
            if not last:
                out.append("(")
            out.append(bond_symbol(b))
            emit(v)
            if not last:
                out.append(")")

    emit(start)
    return "".join(out)


@pytest.mark.parametrize("name", ["cn", "ysi", "lhv"])
def test_perception_does_not_depend_on_how_the_smiles_is_written(name):
    """Every stereo-free molecule of a data set, rewritten from a random atom with random branch order and aromatic
    rings in lower case (so the rewritten string goes through kekulisation), must give the same multiset of atom and
    bond strings: valence, rings, aromaticity, conjugation and hybridisation are properties of the graph."""
    import numpy as np
    rng = np.random.default_rng(7)
    smiles = load_fuel_set(name)[0]
    checked = 0
    for smi in smiles:
        if any(c in smi for c in "@/\\"):
            continue
        mol = Chem.MolFromSmiles(smi)
        if any(a.GetIsAromatic() and a.GetSymbol() == "N" and a.GetTotalNumHs() for a in mol.GetAtoms()):
            continue                              # [nH] must be written in brackets: explicit-H fields differ by design
        want = _strings(smi)
        for _ in range(2):
            again = _write_random_smiles(mol, rng)
            assert _strings(again) == want, (smi, again)
        checked += 1
    assert checked > 250


def test_inputs_outside_the_restated_rules_fail_loudly_or_follow_rdkit():
    assert Chem.MolFromSmiles("[H]C([H])([H])O") is None          # explicit H atoms: RDKit's removeHs is not restated
    assert Chem.MolFromSmiles("[H][H]") is None and Chem.MolFromSmiles("[Fe]") is None
    assert Chem.MolFromSmiles("CCO ethanol").GetNumAtoms() == 3   # "SMILES title"
    pyr = Chem.MolFromSmiles("c1cc[o+]cc1")                       # pyrylium kekulises and is aromatic again
    assert pyr is not None and all(a.GetIsAromatic() for a in pyr.GetAtoms())
    assert [a.GetFormalCharge() for a in pyr.GetAtoms()].count(1) == 1
    assert str(Chem.MolFromSmiles("C[N@](CC)CCC").GetAtomWithIdx(1).GetChiralTag()) == "CHI_UNSPECIFIED"
    with pytest.raises(ValueError):
        MoleculeEncoder(["[H]C"], backend="builtin")
