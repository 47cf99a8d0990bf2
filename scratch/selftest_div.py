import ctypes as C, sys, time
sys.path.insert(0, '.')
import __graft_entry__ as e
dll = C.CDLL(e.LIB)
f = dll.imc_cuda_selftest_div
f.argtypes = [C.c_int, C.c_uint64, C.c_longlong, C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong), C.POINTER(C.c_float)]
per = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
bad, n = C.c_ulonglong(), C.c_ulonglong(); fb = (C.c_float * 4)()
t = time.time(); rc = f(0, 12345, per, C.byref(bad), C.byref(n), fb)
print("rc", rc, "tested %.3e mismatches %d  first_bad a=%r b=%r got=%r want=%r  %.1fs" % (n.value, bad.value, fb[0], fb[1], fb[2], fb[3], time.time() - t))
