cd $GRAFT_REPO_ROOT
for sc in 1 2 1.5; do
 echo "== scale $sc"
 POCO_B200_SHARE_SCALE=$sc timeout 300 python bench.py --steps 10 --no-cpu-baseline --no-other-mode 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cliff fp16', d['value'], d['ms_per_step'])"
 POCO_B200_SHARE_SCALE=$sc timeout 300 python bench.py --steps 6 --precision split --no-cpu-baseline --no-other-mode 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cliff split', d['value'], d['ms_per_step'])"
 POCO_B200_SHARE_SCALE=$sc timeout 300 python bench.py --steps 10 --preset pare_w32 --batch 128 --no-cpu-baseline --no-other-mode 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('pare_w32 fp16', d['value'], d['ms_per_step'])"
 POCO_B200_SHARE_SCALE=$sc timeout 300 python bench.py --steps 10 --preset cliff_w48cls --no-cpu-baseline --no-other-mode 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('w48cls fp16', d['value'], d['ms_per_step'])"
done
