fmt='import sys,json; [print(d["robot"],d["mode"],d["subproblems"],round(d["ms"],3),"ms",round(d["subproblems_per_s"]/1e6,2),"M/s") for d in map(json.loads,sys.stdin)]'
for rep in 1 2; do
echo "== 512 threads (default)"; python scripts/gpu_configs.py 2>/dev/null | python -c "$fmt" | grep centroidal
echo "== 384 threads"; CIMPC_B200_LIB=contactimplicitmpc.jl_b200/lib/libcimpc_b200_g384.so python scripts/gpu_configs.py 2>/dev/null | python -c "$fmt" | grep centroidal
done
