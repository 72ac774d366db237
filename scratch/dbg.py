import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, 'tests')
from hicpeaks_b200 import _capi
from hicpeaks_b200.synth import synth_chromosome
from helpers import compare_with_oracle
generic = sys.argv[1] == 'generic'
inp = synth_chromosome(600, 60, 5, maxww=8, seed=1, scale=300.0)
with _capi.Context(0) as ctx:
    print(compare_with_oracle(ctx, inp, [2], [5], 8, 0.1, 60, 16, generic_kernel=generic))
