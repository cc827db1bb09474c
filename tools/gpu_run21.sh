#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/r2_gputests_21.log; cat gpurun_out/r2_gputests_21.log
timeout 600 python tools/selfcheck_reference_fixtures.py -m gpu -q 2>&1 | tail -4
timeout 600 python tools/bench_modes.py --only c3 2>/dev/null | cut -c1-260
