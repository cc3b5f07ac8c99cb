"""The reference's side scripts that call the same boundary, on the GPU path:
scripts/quantifyLoops.py and scripts/deLoops (the range-count consumers).  Same command lines, same
output files; the per-loop Python loops over getCounts sets became one batched kernel call per chromosome."""
