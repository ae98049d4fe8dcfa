#!/bin/bash
# Wall-clock of an SPR search: the reference's own CPU binary vs the same reference driving the B200
# engine (integration/_build/phyml_b200).  Run on the GPU box:  bash tools/spr_compare.sh 30 20000
NT=${1:-30}; NS=${2:-20000}; ONLY=${3:-both}   # third argument "b200" skips the (slow) CPU reference run
ROOT=$(cd "$(dirname "$0")/.." && pwd)
W=$(mktemp -d)
python - <<PY
import sys; sys.path.insert(0, "$ROOT")
from phyml_b200 import alignment, model as pmodel
from phyml_b200.tree import Tree
tree = Tree.random($NT, seed=5)
m = pmodel.gtr(alpha=0.5)
codes = alignment.simulate(tree, m, $NS, seed=6)
alignment.write_phylip("$W/a.phy", codes, 4, tree.names)
PY
cp $W/a.phy $W/b.phy
cd $W
ARGS="-d nt -m GTR -c 4 -a 0.5 -f e -o tlr -s SPR -b 0 --r_seed 1 --no_memory_check"
t0=$(date +%s.%N)
PLK_SHIM_VERBOSE=1 $ROOT/integration/_build/phyml_b200 -i b.phy $ARGS > b.log 2>&1
t1=$(date +%s.%N)
if [ "$ONLY" != "b200" ]; then $ROOT/oracle/_ref/phyml_ref -i a.phy $ARGS > a.log 2>&1; else echo "(CPU run skipped)" > a.log; fi
t2=$(date +%s.%N)
echo "config: $NT taxa x $NS sites GTR+G4, -o tlr -s SPR"
echo "B200 : wall $(python -c "print(round($t1-$t0,2))") s  $(grep -E 'Log likelihood of the current' b.log | tail -1)  $(grep -E 'Time used' b.log | tail -1)"
grep -E "phyml_b200: (Lk|Pars|wall-clock|.*instance for)" b.log | tail -4
echo "CPU  : wall $(python -c "print(round($t2-$t1,2))") s  $(grep -E 'Log likelihood of the current' a.log | tail -1)  $(grep -E 'Time used' a.log | tail -1)"
