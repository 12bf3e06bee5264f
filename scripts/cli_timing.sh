#!/bin/bash
# phase times of pbk_assemble on C1 (PBK_TIMING=1); usage: scripts/cli_timing.sh [scale]
cd "$(dirname "$0")/.."
D=$(mktemp -d -p /dev/shm pbkt_XXXX)
python - "$D" "${1:-1}" <<'PY'
import sys
sys.path.insert(0, ".")
from platanus_b_b200 import synth
rs = synth.make_reads(synth.config("C1", scale=float(sys.argv[2])))
synth.write_fastq(rs, sys.argv[1] + "/r_1.fq", sys.argv[1] + "/r_2.fq")
PY
ls -la "$D"
for i in 1 2; do
  S=$(date +%s.%N)
  PBK_TIMING=1 platanus_b_b200/_lib/pbk_assemble assemble -kmer_occ_only -k 32 -t 2 -m 16 -tmp "$D" -o "$D/gpu" -f "$D/r_1.fq" "$D/r_2.fq" 2>&1 | grep -E "pbk_assemble|Error"
  E=$(date +%s.%N)
  echo "process wall $(echo "$E - $S" | bc -l 2>/dev/null || python -c "print($E - $S)") s"
done
rm -rf "$D"
