for cs in 30 12 6 3; do for ms in 4 8 16; do
  echo -n "n=256 coarse_sweeps=$cs min_size=$ms: "; timeout 300 python tools/run_cavity.py --n 256 --re 1000 --opt mg_coarse_sweeps=$cs mg_min_size=$ms 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['seconds'], d['krylov_its'], d['newton_its'], d['steps'])"
done; done
for cs in 30 8 4; do for ms in 4 16; do
  echo -n "n=1024 coarse_sweeps=$cs min_size=$ms: "; timeout 300 python tools/run_cavity.py --n 1024 --re 1000 --opt mg_coarse_sweeps=$cs mg_min_size=$ms 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['seconds'], d['krylov_its'], d['newton_its'], d['steps'])"
done; done
