B=/root/repo/e_alphazero_b200/csrc/build
nvidia-smi --query-gpu=serial --format=csv,noheader
echo OLD; EAZ_LIB_PATH=$B/libeaz_base.so timeout 200 python profiles/stress_tensor_subleq.py 250 2>&1 | tail -1
echo NEW; timeout 200 python profiles/stress_tensor_subleq.py 250 2>&1 | tail -1
echo OLD; EAZ_LIB_PATH=$B/libeaz_base.so timeout 200 python profiles/stress_tensor_subleq.py 250 2>&1 | tail -1
echo NEW; timeout 200 python profiles/stress_tensor_subleq.py 250 2>&1 | tail -1
