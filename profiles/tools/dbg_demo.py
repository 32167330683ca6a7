import sys, os
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import helpers, oracle_binding as ob
from longcallr_b200 import abi, host
reads, refs, regions = helpers.load_demo_fixture()
p = host.params_preset("hifi-masseq", seed=7, flags=abi.LCR_FLAG_EMIT_PLANES | abi.LCR_FLAG_EMIT_FRAGMENTS)
batch = host.BatchView(reads, regions)
eng = host.Engine(p, device=0)
eng.set_references(refs)
try:
    got = eng.submit(batch)
    want = ob.run(p, batch, refs, mode=0)
    helpers.compare_results(got, want, "demo")
    print("demo ok")
except Exception as e:
    print("ERR", e)
