# Development sweep (run under gpurun): L2 eviction policies x tile configurations of the fused H psi kernel;
# produced gpurun_out/r02_try2.log, summarised in profiles/r02_hpsi_ncu.md.
for pol in 0 1 2; do
 echo "== POL $pol"
 MGB_HPSI_POL=$pol python tools/cfg_try.py --n 256 --orb 256 --dtype f64 --lap 2 4,2,1,4,64:0 2,4,1,4,64:0 4,2,1,3,128:0 4,2,1,4,0:0 | cut -c1-130
 MGB_HPSI_POL=$pol python tools/cfg_try.py --n 256 --orb 256 --dtype f64 --lap 0 8,2,1,3,128:0 8,2,1,3,0:0| cut -c1-130
 MGB_HPSI_POL=$pol python tools/cfg_try.py --n 128 --orb 256 --dtype f64 --lap 2 | cut -c1-130
 MGB_HPSI_POL=$pol python tools/cfg_try.py --n 128 --orb 256 --dtype f64 --lap 0 | cut -c1-130
done
