#!/bin/bash
# GPU box: duration reported by the drop-in CLI for configs 1 and 2 (121 grid points each), lockstep off / on
BIN=$GRAFT_REPO_ROOT/bose-hubbard-phase-transition_b200/QuantumProject
for m in 8 10; do for b in 1 4; do
  d=/tmp/cli_${m}_$b; mkdir -p $d; cd $d
  $BIN -t exact -m $m -n $m -J 1 -U 0 -u 0 -r 10 -s 1 -f J --kernel free --batch $b --no-plot > out.log 2>&1
  echo "m=$m batch=$b: $(tr '\r' '\n' < out.log | grep 'Calculation duration') phase.txt md5 $(md5sum phase.txt | cut -c1-12) rows $(wc -l < phase.txt)"
done; done
