# forward rollout at fewer resident blocks per SM, run-to-completion schedule (RLSDE_FWD_QUANTUM=0)
export RLSDE_FWD_QUANTUM=0
for b in 1 2 4 8; do echo -n "bps=$b "; RLSDE_FWD_BLOCKS_PER_SM=$b python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'])"; done
