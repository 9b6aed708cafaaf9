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


def write_input(d):
    from isonclust2_b200 import synth
    seq, qual, offs = synth.make_reads(N_READS, seed=SEED, reads_per_gene=READS_PER_GENE)
    synth.write_fastq(os.path.join(d, "reads.fq"), seq, qual, offs)


def digest(path):
    d = open(path, "rb").read()
    return hashlib.sha256(d[:28] + b"\0\0\0\0" + d[32:]).hexdigest()


def run_pipeline(exe, d, steps=STEPS, timeout=1500):
    """Runs sort + every step in directory d; returns {file: digest} and the stderr of the verbose last merge."""
    out = {}
    r = subprocess.run([exe] + SORT_ARGS, cwd=d, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=timeout)
    assert r.returncode == 0, r.stderr[-2000:]
    for i in range(3):
        out[f"isONbatch_{i}.cer"] = digest(os.path.join(d, "srt", "batches", f"isONbatch_{i}.cer"))
    for name, args in steps:
        r = subprocess.run([exe] + args, cwd=d, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=timeout)
        assert r.returncode == 0, (args, r.stderr[-2000:])
        out[name] = digest(os.path.join(d, name))
    return out
