"""Feasibility probe: does CUDA IPC (cudaIpcGetMemHandle / cudaIpcOpenMemHandle) work between two
processes on two GPUs of this box, and what does a peer copy through the mapping reach?"""
import ctypes as C, multiprocessing as mp, time, sys
def rt():
    for nm in ("libcudart.so.12", "/usr/local/cuda/lib64/libcudart.so"):
        try:
            return C.CDLL(nm)
        except OSError:
            pass
    raise SystemExit("no libcudart")
def child(conn):
    L = rt()
    assert L.cudaSetDevice(1) == 0
    h = conn.recv()
    hb = (C.c_ubyte * 64).from_buffer_copy(h)
    p = C.c_void_p()
    rc = L.cudaIpcOpenMemHandle(C.byref(p), hb, 1)   # cudaIpcMemLazyEnablePeerAccess; struct passed by value below
    conn.send(("open", rc))
def main():
    L = rt()
    class H(C.Structure):
        _fields_ = [("r", C.c_ubyte * 64)]
    L.cudaIpcOpenMemHandle.argtypes = [C.POINTER(C.c_void_p), H, C.c_uint]
    L.cudaIpcGetMemHandle.argtypes = [C.POINTER(H), C.c_void_p]
    n = 1 << 28
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        return
    ctx = mp.get_context("spawn")
    a, b = ctx.Pipe()
    pr = ctx.Process(target=child2, args=(b, n))
    pr.start()
    assert L.cudaSetDevice(0) == 0
    d = C.c_void_p()
    assert L.cudaMalloc(C.byref(d), C.c_size_t(n)) == 0
    L.cudaMemset(d, 0, C.c_size_t(n))
    h = H()
    rc = L.cudaIpcGetMemHandle(C.byref(h), d)
    print("get handle rc", rc, flush=True)
    a.send(bytes(h.r))
    print("child:", a.recv(), flush=True)
    print("child:", a.recv(), flush=True)
    buf = (C.c_ubyte * 16)()
    L.cudaMemcpy(buf, d, 16, 2)
    print("first bytes after peer write:", list(buf), flush=True)
    pr.join()
def child2(conn, n):
    L = rt()
    class H(C.Structure):
        _fields_ = [("r", C.c_ubyte * 64)]
    L.cudaIpcOpenMemHandle.argtypes = [C.POINTER(C.c_void_p), H, C.c_uint]
    assert L.cudaSetDevice(1) == 0
    hb = conn.recv()
    h = H()
    C.memmove(h.r, hb, 64)
    p = C.c_void_p()
    rc = L.cudaIpcOpenMemHandle(C.byref(p), h, 1)
    conn.send(("open rc", rc))
    if rc != 0:
        conn.send(("skip", 0)); return
    src = C.c_void_p()
    L.cudaMalloc(C.byref(src), C.c_size_t(n))
    L.cudaMemset(src, 7, C.c_size_t(n))
    L.cudaDeviceSynchronize()
    e0, e1 = C.c_void_p(), C.c_void_p()
    L.cudaEventCreate(C.byref(e0)); L.cudaEventCreate(C.byref(e1))
    L.cudaMemcpy(p, src, C.c_size_t(n), 3)
    L.cudaEventRecord(e0, None)
    for _ in range(5):
        L.cudaMemcpyAsync(p, src, C.c_size_t(n), 3, None)
    L.cudaEventRecord(e1, None)
    L.cudaDeviceSynchronize()
    ms = C.c_float()
    L.cudaEventElapsedTime(C.byref(ms), e0, e1)
    conn.send(("peer memcpy GB/s", 5 * n / (ms.value * 1e-3) / 1e9))
if __name__ == "__main__":
    main()
