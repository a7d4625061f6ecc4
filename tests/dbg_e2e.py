import sys, os, subprocess
sys.path.insert(0, 'tests')
import synth, oracle_lib as O
cfg=dict(seed=3, glen=500_000, n=300, rlen=3000, err=0.05, rep=0.1, chim=0.05)
contigs = synth.random_genome(cfg["glen"], cfg["seed"], n_contigs=3, repeat_frac=cfg["rep"])
reads = synth.make_reads(contigs, cfg["n"], cfg["rlen"], cfg["err"], cfg["seed"] + 100, chimeric_frac=cfg["chim"])
os.makedirs('gpurun_out', exist_ok=True)
synth.write_fasta('gpurun_out/ref.fa', ["chr%d" % i for i in range(len(contigs))], contigs)
synth.write_fasta('gpurun_out/reads.fa', ["read%d" % i for i in range(len(reads))], reads)
a=subprocess.run([O.REF_BIN,'-t4','-x','map-ont','-c','gpurun_out/ref.fa','gpurun_out/reads.fa'],stdout=subprocess.PIPE,stderr=subprocess.DEVNULL).stdout.decode().splitlines()
b=subprocess.run(['minimap2_b200/minimap2-b200','-t8','-x','map-ont','-c','gpurun_out/ref.fa','gpurun_out/reads.fa'],stdout=subprocess.PIPE,stderr=subprocess.DEVNULL).stdout.decode().splitlines()
open('gpurun_out/ref.paf','w').write("\n".join(a)); open('gpurun_out/got.paf','w').write("\n".join(b))
import difflib
for l in difflib.unified_diff([x[:200] for x in a],[x[:200] for x in b],lineterm='',n=0): print(l)
