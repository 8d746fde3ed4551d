set -x
mkdir -p gpurun_out
for cfg in "131072 2" "262144 2" "262144 3" "131072 3" "65536 3"; do set -- $cfg; python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-cnn --no-dense --block-ligands $1 --slots $2 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('block $1 slots $2: value %.2f M  e2e %.2f M (%.1f ms)'%(d['value']/1e6, d['e2e']['value']/1e6, d['e2e']['ms_per_step']))"; done
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-cnn --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps(d['dense_model']))"
