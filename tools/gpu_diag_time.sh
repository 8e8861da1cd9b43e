#!/bin/bash
mkdir -p gpurun_out
date +%T
( time timeout 120 python -c "import torch; print(torch.zeros(1).cuda())" ) 2>&1 | tail -4
date +%T
( time timeout 300 python -X faulthandler -c "
import faulthandler, sys, time
faulthandler.dump_traceback_later(100, exit=True)
t0=time.time()
sys.argv=['x','2']
exec(open('tools/prof_frame.py').read())
print('script done', time.time()-t0, flush=True)
" ) 2>&1 | tail -30
date +%T
