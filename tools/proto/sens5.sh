for opts in "ksp_rtol=1e-5" "ksp_rtol=1e-4" "ksp_rtol=1e-3" "ksp_rtol=1e-2"; do
  for n in 256 1024; do echo -n "n=$n $opts: "; timeout 200 python tools/run_cavity.py --n $n --re 1000 --opt $opts 2>&1 | tail -1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print(d['seconds'], 'krylov', d['krylov_its'], 'newton', d['newton_its'], 'steps', d['steps'], 'du', d['final_du'])
except Exception as e: print('FAILED')"
done; done
