"""RECORD ONLY (profiles/r2_ab_multi_body_edges.jsonl): cost of the edge-through-box contacts and of the separating-axis pre-test
in step_multi_body_kernel, measured with a temporary launch flag that skipped them (bit 0 = no edge contacts, bit 1 = no
pre-test; it changed the results and was removed from the library again, so this script no longer runs)."""
import json, sys, os
sys.path.insert(0,'.')
import torch
import rigidbody_simulation_b200 as rb
from rigidbody_simulation_b200 import stepper
from rigidbody_simulation_b200.src.simulation import mixed_pile
E=131072
for skip in (0,1,2,3):
    rb._lib.set_option("mb_debug_skip", skip)
    model, data = mixed_pile.build(E, device="cuda:0", dtype=torch.float64, n_body=8)
    best=None
    for rep in range(2):
        model, data = mixed_pile.build(E, device="cuda:0", dtype=torch.float64, n_body=8)
        ev=[torch.cuda.Event(enable_timing=True) for _ in range(5)]
        for i in range(4):
            ev[i].record(); stepper.step_multi_body(model, data, 0.005, 0.2, 0.6, substeps=64, count=True)
        ev[4].record(); torch.cuda.synchronize()
        ms=[ev[i].elapsed_time(ev[i+1]) for i in range(4)]
        if best is None or sum(ms)<sum(best): best=ms
    c,i=data.counters()
    print(json.dumps({"skip":skip,"ms":[round(m,1) for m in best],"body_substeps_per_s":E*8*256/(sum(best)*1e-3),"contacts_per_body_substep":float(c.sum())/(E*8*256)}))
rb._lib.set_option("mb_debug_skip", 0)
