#!/usr/bin/env python
"""The kernels' text under AddressSanitizer + UBSan (CPU, tests/cuda_emul): payload / output / counts buffers are heap
blocks of EXACTLY the documented sizes, so any access past n (payload), n + 1 (serialised output) or the tile
rectangle trips a red zone; integer UB (shifts, overflow) trips UBSan.

    cd tests/cuda_emul && g++ -O1 -g -std=c++17 -fPIC -shared -ffp-contract=off -fsanitize=address,undefined \
        -fno-sanitize-recover=undefined -w -I. -I../../distributedmandelbrot_b200/csrc -o /tmp/libkernels_asan.so \
        cuda_emul.cpp codec_emul.cpp escape_emul.cpp
    LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0 python scripts/emul_asan.py
"""
import ctypes as C, numpy as np, sys
import os
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT,'oracle'), os.path.join(ROOT,'tests')): sys.path.insert(0,p)
import chunk_codec as codec, oracle
from distributedmandelbrot_b200 import _native as nat
L=C.CDLL('/tmp/libkernels_asan.so')
L.emul_serialize_chunks.argtypes=[C.c_void_p, C.c_size_t, C.c_uint32, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_int, C.c_int, C.c_uint]
L.emul_render_counts.argtypes=[C.c_void_p, C.c_uint32, C.c_void_p, C.c_int, C.c_void_p]
L.emul_render_u8.argtypes=[C.c_void_p, C.c_uint32, C.c_void_p, C.c_int, C.c_void_p]
rng=np.random.default_rng(1)
# serializer: exact-size heap buffers (ASan red zones right behind payload and output)
for n in (1, 17, 8191, 16384, 16385, 70001, 200000):
    for kind in range(4):
        if kind==0: p=rng.integers(0,256,n).astype(np.uint8)
        elif kind==1: p=np.repeat(rng.integers(0,4,n//3+2).astype(np.uint8), rng.integers(1,12,n//3+2))[:n]
        elif kind==2: p=np.full(n,7,np.uint8)
        else:
            p=np.full(n,3,np.uint8); p[n//4:n//2]=rng.integers(0,256,n//2-n//4)
        p=np.ascontiguousarray(p)
        for spans in (1,2):
            out=np.empty(n+1,np.uint8); info=np.zeros(32,np.uint8)     # out_stride == n + 1 exactly
            assert L.emul_serialize_chunks(p.ctypes.data,n,1,n,out.ctypes.data,n+1,info.ctypes.data,1,spans,65535)==0
            assert out[:int(info[:8].view('<u8')[0])].tobytes()==codec.serialize(p)
print("serializer ok")
# escape kernels: ragged tiles, exact-size outputs
for (level,ir,ii,d,mrd) in [(10,3,5,77,300),(1,0,0,129,100),(800,250,420,50,2000)]:
    want=oracle.render_chunk_counts(level,mrd,ir,ii,d)
    for sc in (False,True):
        for rf in (False,True):
            t=nat.chunk_geometry(level,ir,ii,d,mrd,shortcut=sc,refill=rf)
            out=np.empty(d*d,np.int32); ex=C.c_ulonglong(0)
            assert L.emul_render_counts(nat.tile_array([t]),1,out.ctypes.data,0,C.byref(ex))==0
            assert np.array_equal(out.reshape(d,d),want)
            o8=np.empty(d*d,np.uint8)
            assert L.emul_render_u8(nat.tile_array([t]),1,o8.ctypes.data,0,C.byref(ex))==0
            assert np.array_equal(o8.reshape(d,d),oracle.counts_to_u8(want,mrd))
print("escape ok")
