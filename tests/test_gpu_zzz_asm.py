"""GPU end-to-end parity for the assembly presets (MM_F_RMQ: mg_lchain_rmq as the first chainer, map.c:275-276): CLI output vs
the unmodified reference binary, byte for byte. Sorted last: this path was added after the round's last B200 session."""
import os
import numpy as np
import pytest
import oracle_lib as O
import synth
from test_gpu_e2e import compare, DATA

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(not os.path.exists(O.REF_BIN), reason="oracle/_ref not built")
@pytest.mark.parametrize("preset", ["asm5", "asm10", "asm20"])
def test_mt_assembly_presets(preset):
    compare(["-x", preset, "-c", "--cs", os.path.join(DATA, "MT-human.fa"), os.path.join(DATA, "MT-orang.fa")])


@pytest.mark.skipif(not os.path.exists(O.REF_BIN), reason="oracle/_ref not built")
@pytest.mark.parametrize("preset,div", [("asm5", 0.003), ("asm10", 0.02), ("asm20", 0.05)])
def test_synthetic_contigs(tmp_path, preset, div):
    rng = np.random.default_rng(77)
    contigs = synth.random_genome(1_500_000, 31, n_contigs=2, repeat_frac=0.15)
    comp = np.zeros(256, dtype=np.uint8); comp[list(b"ACGT")] = list(b"TGCA")
    asm = []
    for i in range(24):  # contigs of 20-120 kb with a deletion, an insertion and sometimes an inverted segment
        c = np.frombuffer(bytes(contigs[i % 2]), dtype=np.uint8)
        L = int(rng.integers(20_000, 120_000)); s = int(rng.integers(0, len(c) - L))
        x = c[s:s + L]
        d0 = int(rng.integers(2000, L // 2)); dl = int(rng.integers(50, 3000))
        parts = [x[:d0], x[d0 + dl:L * 2 // 3], synth.ALPHA[rng.integers(0, 4, int(rng.integers(30, 1500)))], x[L * 2 // 3:]]
        if i % 3 == 0:
            parts[-1] = comp[parts[-1][::-1]]
        y = np.concatenate(parts)
        if i % 2:
            y = comp[y[::-1]]
        asm.append(synth.mutate_ascii(y, rng, div))
    rf, qf = str(tmp_path / "ref.fa"), str(tmp_path / "asm.fa")
    synth.write_fasta(rf, ["chr%d" % i for i in range(len(contigs))], contigs)
    synth.write_fasta(qf, ["ctg%d" % i for i in range(len(asm))], asm)
    n = compare(["-x", preset, "-c", "--cs", rf, qf])
    assert n >= len(asm)
