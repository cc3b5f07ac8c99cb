"""Throughput of the BEDPE ingest: lines/s of cloops_bedpe_parse against the pandas tokenizer path it
replaced and -- on a prefix -- the reference's own parseRawBedpe2 + txt2jd (cLoops/io.py:132-203, through the shim).
    python tools/ingest_probe.py [lines] [workdir]"""
import gzip
import logging
import os
import shutil
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cloops_b200 import io  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 5_000_000
work = sys.argv[2] if len(sys.argv) > 2 else tempfile.mkdtemp()
rng = np.random.default_rng(3)
chrom = rng.integers(1, 24, n)
a = rng.integers(0, 2 * 10 ** 8, n)
d = rng.integers(0, 10 ** 6, n)
plain = os.path.join(work, "probe.bedpe")
t0 = time.time()
with open(plain, "w") as fh:
    for lo in range(0, n, 500000):
        fh.write("".join("chr%d\t%d\t%d\tchr%d\t%d\t%d\tSRR0000000.%d\t255\t%s\t%s\n" % (c, x, x + 36, c, x + y, x + y + 36, k, "+-"[k & 1], "+-"[(k >> 1) & 1])
                         for k, (c, x, y) in enumerate(zip(chrom[lo:lo + 500000].tolist(), a[lo:lo + 500000].tolist(), d[lo:lo + 500000].tolist()), lo)))
gz = plain + ".gz"
with open(plain, "rb") as src, gzip.open(gz, "wb", compresslevel=6) as dst:
    shutil.copyfileobj(src, dst, 1 << 24)
print("wrote %d lines, %.0f MB plain, %.0f MB gz in %.0f s" % (n, os.path.getsize(plain) / 1e6, os.path.getsize(gz) / 1e6, time.time() - t0))
log = logging.getLogger("probe")
for label, f in (("plain", plain), ("gzip", gz)):
    for threads in (1, 0):
        io.INGEST_THREADS = threads
        dt = 1e9
        for rep in range(2):                                   # best of two: the first call also grows the heap
            t0 = time.time()
            order, per, total = io._cis_native([f], [], 0)
            dt = min(dt, time.time() - t0)
        print("native %-5s threads=%-2s %.2f s  %.2f M lines/s  (%d PETs, %d chromosomes)" % (
            label, threads or os.cpu_count(), dt, total / dt / 1e6, sum(len(per[c][0]) for c in order), len(order)))
    io.INGEST_THREADS = 0
    t0 = time.time()
    out = os.path.join(work, "jd_" + label)
    os.mkdir(out)
    io.parseRawBedpe2([f], out, [], 0, log)
    dt = time.time() - t0
    print("parseRawBedpe2 %-5s (native + .jd write) %.2f s  %.2f M lines/s" % (label, dt, n / dt / 1e6))
    if n <= 20_000_000:
        t0 = time.time()
        io._cis_table(f, [], 0)
        dt = time.time() - t0
        print("pandas tokenizer path %-5s %.2f s  %.2f M lines/s" % (label, dt, n / dt / 1e6))
try:
    from oracle import ref_shim
    if ref_shim.available():
        ns = ref_shim.load()
        m = min(n, 300000)
        head = os.path.join(work, "head.bedpe")
        with open(plain) as src, open(head, "w") as dst:
            for k, line in enumerate(src):
                if k >= m:
                    break
                dst.write(line)
        out = os.path.join(work, "ref")
        os.mkdir(out)
        t0 = time.time()
        for f in ns.io.parseRawBedpe2([head], out, [], 0, log):
            ns.io.txt2jd(f)
        dt = time.time() - t0
        print("reference parseRawBedpe2 + txt2jd on the first %d lines: %.2f s  %.3f M lines/s" % (m, dt, m / dt / 1e6))
except Exception as e:      # the probe is a measurement aid
    print("reference leg skipped:", e)
shutil.rmtree(work, ignore_errors=True)
