# round 2, GPU call P: which switch removes the run-to-run difference of the step-1 loss
mkdir -p gpurun_out
{
timeout 200 python tests/repro_fit_race.py 2>&1 | tail -1
HALO=0 timeout 200 python tests/repro_fit_race.py 2>&1 | tail -1
PERSIST=0 timeout 200 python tests/repro_fit_race.py 2>&1 | tail -1
HALO=0 PERSIST=0 timeout 200 python tests/repro_fit_race.py 2>&1 | tail -1
PDL=0 timeout 200 python tests/repro_fit_race.py 2>&1 | tail -1
GRAPH=0 timeout 200 python tests/repro_fit_race.py 2>&1 | tail -1
OVERLAP=0 timeout 200 python tests/repro_fit_race.py 2>&1 | tail -1
} | grep -v "^Epoch" | tee gpurun_out/repro_fit_race_r2p.log
