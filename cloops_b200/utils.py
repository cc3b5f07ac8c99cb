"""Logger and command line of the ``cLoops`` entry point (cLoops/utils.py:23-204): same flags, same
defaults; ``-p`` keeps its meaning of "workers", which here are GPUs (one process per GPU)."""
import argparse
import logging
import sys
import time

__version__ = "0.93-b200"
EPILOG = "B200-native drop-in for the cLoops clustering and scoring path."


def getLogger(fn=None):
    """cLoops/utils.py:23-44: INFO log to ``fn`` and to stdout.  Created once per process (the
    reference appends a new stdout handler on every call)."""
    logging.basicConfig(level=logging.INFO, format="%(asctime)s %(name)-6s %(levelname)-8s %(message)s",
                        datefmt="%Y-%m-%d %H:%M:%S", filename=fn, filemode="a")
    logger = logging.getLogger()
    if not any(getattr(h, "_cloops_stdout", False) for h in logger.handlers):
        handler = logging.StreamHandler(sys.stdout)
        handler.setFormatter(logging.Formatter("%(asctime)s %(levelname)s %(message)s"))
        handler._cloops_stdout = True
        logger.addHandler(handler)
    logger.setLevel(logging.NOTSET)
    return logger


def cFlush(r):
    """cLoops/utils.py:65-70."""
    sys.stdout.write("\r%s" % r)
    sys.stdout.flush()


def mainHelp(argv=None):
    """cLoops/utils.py:73-204: the ``cLoops`` flags -f -o -m -eps -minPts -p -c -w -j -s -hic -cut -max_cut -plot -v."""
    p = argparse.ArgumentParser(description="Intra-chromosomal loops calling for ChIA-PET,HiChIP and high-resolution Hi-C data.",
                                epilog=EPILOG)
    p.add_argument("-f", dest="fnIn", required=True, type=str,
                   help="Mapped PETs, BEDPE (optionally gzipped); replicates as A.bedpe.gz,B.bedpe.gz are pooled.")
    p.add_argument("-o", dest="fnOut", required=True, type=str, help="Output prefix.")
    p.add_argument("-m", dest="mode", required=False, type=int, default=0, choices=[0, 1, 2, 3, 4],
                   help="Pre-set parameters: 0 use -eps/-minPts; 1 sharp-peak ChIA-PET; 2 broad-peak ChIA-PET; 3 deep Hi-C; 4 HiChIP.")
    p.add_argument("-eps", dest="eps", default=0, required=False,
                   help="DBSCAN eps, one value or a comma list (1000,2000); 0 = estimate from the data.")
    p.add_argument("-minPts", dest="minPts", default=0, help="DBSCAN minPts, one value or a comma list.")
    p.add_argument("-p", dest="cpu", required=False, default=1, type=int,
                   help="Workers. In this build a worker is a GPU (launch with torchrun for more than one).")
    p.add_argument("-c", dest="chroms", required=False, default="", type=str, help="Restrict to chr1,chr2,...")
    p.add_argument("-w", dest="washU", required=False, action="store_true", help="Also write a washU long-range track.")
    p.add_argument("-j", dest="juice", required=False, action="store_true", help="Also write Juicebox 2D annotations.")
    p.add_argument("-s", dest="tmp", required=False, action="store_true", help="Keep the per-chromosome .jd directory.")
    p.add_argument("-hic", dest="hic", required=False, action="store_true", help="HiChIP / Hi-C significance cut-offs.")
    p.add_argument("-cut", dest="cut", required=False, default=0, type=int, help="Initial distance cut-off (debugging).")
    p.add_argument("-max_cut", dest="max_cut", required=False, action="store_true",
                   help="Use the largest estimated self/inter-ligation cut-off instead of the smallest.")
    p.add_argument("-plot", dest="plot", required=False, action="store_true",
                   help="Accepted for compatibility; plotting is not part of this build.")
    p.add_argument("-v", dest="version", action="version", version="cLoops v%s" % __version__)
    return p.parse_args(argv)
