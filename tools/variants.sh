SLU_BOTH=1 SLU_YARD=1 python tools/time_reduce.py
for f in semanticlidarunc_b200/libslu_t*.so; do SLU_LIB_PATH=$PWD/$f python tools/time_reduce.py; done
