O=gpurun_out; mkdir -p $O
cat > /tmp/san_case.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np
from util import GAMMA, RGAS, make_state, product_flux, product_setup
import spade_b200.api as sp
for nb, n in (((2, 1, 2), (40, 12, 8)), ((2, 2, 1), (16, 16, 8)), ((1, 2, 2), (12, 20, 6))):
    _, blocks, grid = product_setup(nb, n, 2)
    q0 = make_state(nb, n, 2, seed=3)
    gas = sp.ideal_gas_t(GAMMA, RGAS)
    flux = sp.flux_desc(product_flux(0))
    qa, ra = sp.grid_array.from_host(grid, q0), sp.grid_array(grid, 0.0)
    ex = sp.make_exchange(qa, (1, 1, 1))
    ti = sp.integrator_t(sp.time_axis_t(0.0, 1e-6), sp.rk4_t, sp.integrator_data_t(qa, ra, sp.rk4_t), sp.flux_div_rhs_t(flux, sp.overwrite),
                         sp.exchange_bc_t(ex), sp.state_transform_t(gas))
    ti.advance()
    sp.flux_div(qa, ra, flux, sp.overwrite)
    sp.flux_div(qa, ra, sp.flux_desc(product_flux(1)), sp.increment)
    print("ok", nb, n, float(ti.solution().data.abs().max()))
PY
for tool in racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 5 python /tmp/san_case.py > $O/sanitizer_$tool.log 2>&1; echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|^ok|Error|hazard" $O/sanitizer_$tool.log | head -12
done
