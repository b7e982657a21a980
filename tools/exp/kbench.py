import ctypes as C, sys, os
for name in sys.argv[1:]:
    L = C.CDLL(os.path.abspath(name))
    ctx = C.c_void_p(); assert L.s252_ctx_create(0, C.byref(ctx)) == 0
    k = C.c_double(); L.s252_microbench_keccak(ctx, C.byref(k)); print(name, 'keccak Gperm/s', k.value)
