"""Write the first N column blocks of the 500-taxon x 1M-site workload (phyml_b200/workloads.py) as c4.phy + c4.nwk in the
current directory: input of the drop-in binary for a configs[3]-sized run.  usage: python tools/gen_config4_phylip.py N"""
import sys, time
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from phyml_b200 import alignment, workloads as wl
w = wl.WORKLOADS["dna_500x1M"]
nb = int(sys.argv[1])
tree = wl.make_tree(w)
t=time.time()
codes = np.concatenate([wl.block_codes(w, b) for b in range(nb)], axis=1)
alignment.write_phylip("c4.phy", codes, 4, tree.names)
open("c4.nwk","w").write(tree.to_newick()+"\n")
print("generated", codes.shape, time.time()-t)
