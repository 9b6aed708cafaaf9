"""The consensus-on pipeline both CLIs are run through (test infrastructure): `sort -c <positive>` of a small FASTQ, `cluster` of
every batch and the merges ((0,1),2); every `.cer` is digested with Batch::TotalReads masked (cerfile.py)."""
import hashlib
import os
import subprocess

SORT_ARGS = ["sort", "-B", "50000", "-M", "80", "-c", "12", "-g", "4", "-o", "srt", "reads.fq"]
N_READS, SEED, READS_PER_GENE = 240, 5, 30
# (output, arguments): spoa semi-global (the default, args.h:53), then local and global on one batch
STEPS = [("c0.cer", ["cluster", "-x", "sahlin", "-l", "srt/batches/isONbatch_0.cer", "-o", "c0.cer"]),
         ("c1.cer", ["cluster", "-x", "sahlin", "-l", "srt/batches/isONbatch_1.cer", "-o", "c1.cer"]),
         ("c2.cer", ["cluster", "-x", "fast", "-l", "srt/batches/isONbatch_2.cer", "-o", "c2.cer"]),
         ("m01.cer", ["cluster", "-x", "sahlin", "-l", "c0.cer", "-r", "c1.cer", "-o", "m01.cer"]),
         ("m012.cer", ["cluster", "-x", "sahlin", "-l", "m01.cer", "-r", "c2.cer", "-o", "m012.cer"]),
         ("a0.cer", ["cluster", "-x", "sahlin", "-A", "0", "-l", "srt/batches/isONbatch_0.cer", "-o", "a0.cer"]),
         ("a1.cer", ["cluster", "-x", "sahlin", "-A", "1", "-l", "srt/batches/isONbatch_0.cer", "-o", "a1.cer"])]


# second case: one batch of four large clusters with the whole life of a consensus graph — 15 sequences before the first
# consensus (ConsMinSize: sub-batches of several reads), restarted above 40 (ConsPurge), abandoned above 60 reads (ConsPeriod)
P_SORT_ARGS = ["sort", "-B", "50000", "-M", "300", "-c", "40", "-g", "15", "-P", "60", "-o", "srt", "reads.fq"]
P_READS = dict(n_reads=300, seed=9, reads_per_gene=75, max_len=1200)
P_STEPS = [("p0.cer", ["cluster", "-x", "sahlin", "-l", "srt/batches/isONbatch_0.cer", "-o", "p0.cer"])]


def write_input(d, reads=None):
    from isonclust2_b200 import synth
    kw = reads or dict(n_reads=N_READS, seed=SEED, reads_per_gene=READS_PER_GENE)
    seq, qual, offs = synth.make_reads(kw.pop("n_reads"), **kw)
    synth.write_fastq(os.path.join(d, "reads.fq"), seq, qual, offs)


def digest(path):
    d = open(path, "rb").read()
    return hashlib.sha256(d[:28] + b"\0\0\0\0" + d[32:]).hexdigest()


def run_pipeline(exe, d, steps=STEPS, timeout=1500, sort_args=SORT_ARGS, n_batches=3):
    """Runs sort + every step in directory d; returns {file: digest}."""
    out = {}
    r = subprocess.run([exe] + sort_args, cwd=d, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=timeout)
    assert r.returncode == 0, r.stderr[-2000:]
    for i in range(n_batches):
        out[f"isONbatch_{i}.cer"] = digest(os.path.join(d, "srt", "batches", f"isONbatch_{i}.cer"))
    for name, args in steps:
        r = subprocess.run([exe] + args, cwd=d, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=timeout)
        assert r.returncode == 0, (args, r.stderr[-2000:])
        out[name] = digest(os.path.join(d, name))
    return out
