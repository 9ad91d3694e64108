B="mg_coarse_sweeps=10 gmres_refine=0"
for opts in "$B gmres_restart=50 pc_omega_u=0.8 pc_omega_p=0.8" "$B gmres_restart=60 pc_omega_u=0.8 pc_omega_p=0.8" "$B gmres_restart=50 pc_omega_u=0.9 pc_omega_p=0.9" "$B gmres_restart=50 pc_omega_u=0.9 pc_omega_p=0.7" "$B gmres_restart=50 pc_omega_u=0.7 pc_omega_p=0.9" "$B gmres_restart=50 pc_omega_u=1.0 pc_omega_p=0.8"; do
  echo -n "n=1024 $opts: "; timeout 300 python tools/run_cavity.py --n 1024 --re 1000 --opt $opts 2>&1 | tail -1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print(d['seconds'], d['krylov_its'], d['newton_its'], d['steps'])
except Exception as e: print('FAILED')"
done
