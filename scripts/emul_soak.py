#!/usr/bin/env python
"""Random soak of the kernels' text on the CPU emulator (tests/cuda_emul) against the oracle / the restated serializer.

    python scripts/emul_soak.py escape SEED SECONDS     # random levels, chunks near the set, sub-rectangles, pitches, budgets around the
                                                        # block boundaries, f64 / f32, int32 / uint8, all four kernel variants
    python scripts/emul_soak.py codec SEED SECONDS      # random run structures, batch sizes, alignments, tile geometries, look-back modes,
                                                        # grid shapes and warp interleavings
Stops at the first mismatch (printing the case).  Record: profiles/r02_emulated_soak.txt."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, _p)
MODE = sys.argv[1]
sys.argv = [sys.argv[0]] + sys.argv[2:]


def escape_soak():
    import ctypes as C, time
    import numpy as np
    from distributedmandelbrot_b200 import _native as nat
    import oracle
    from test_escape_emulated import build_emulator
    L=C.CDLL(build_emulator())
    for fn in (L.emul_render_counts, L.emul_render_u8): fn.argtypes=[C.c_void_p, C.c_uint32, C.c_void_p, C.c_int, C.c_void_p]
    rng=np.random.default_rng(int(sys.argv[1]) if len(sys.argv)>1 else 1)
    t_end=time.time()+float(sys.argv[2]) if len(sys.argv)>2 else time.time()+600
    cases=0; pix=0
    while time.time()<t_end:
        mode=rng.integers(0,4)
        if mode==0: level=int(rng.integers(1,40))
        elif mode==1: level=int(rng.integers(40,5000))
        elif mode==2: level=int(rng.integers(5000,5_000_000))
        else: level=int(rng.choice([1,2,4,5,8,10,16,20,800,4000000]))
        # bias towards chunks that touch the set: pick a point near the boundary and take its chunk
        th=rng.uniform(0,2*np.pi); 
        pr=0.5*np.cos(th)-0.25*np.cos(2*th)+rng.normal(0,0.02); pi=0.5*np.sin(th)-0.25*np.sin(2*th)+rng.normal(0,0.02)
        if rng.random()<0.3: pr,pi=-1+0.25*np.cos(th)+rng.normal(0,0.01),0.25*np.sin(th)+rng.normal(0,0.01)
        ir=int(np.clip((pr+2)/4*level,0,level-1)); ii=int(np.clip((pi+2)/4*level,0,level-1))
        d=int(rng.integers(2,140)); mrd=int(rng.choice([1,2,3,7,8,9,16,17,31,32,33,34,40,41,47,48,49,63,64,65,100,255,256,1000,3000]))
        w=int(rng.integers(1,d+1)); h=int(rng.integers(1,d+1)); x0=int(rng.integers(0,d-w+1)); y0=int(rng.integers(0,d-h+1))
        pitch=w+int(rng.integers(0,5)); off=int(rng.integers(0,7))
        dtype=str(rng.choice(["f64","f64","f32"])); sc=bool(rng.integers(0,2)); rf=bool(rng.integers(0,2)); u8=bool(rng.integers(0,2))
        chunk=nat.chunk_geometry(level,ir,ii,d,mrd,shortcut=sc,refill=rf)
        t=chunk.sub(x0,y0,w,h,off,pitch)
        n_out=off+pitch*h+8
        out=np.full(n_out,77 if u8 else -7,np.uint8 if u8 else np.int32)
        fn=L.emul_render_u8 if u8 else L.emul_render_counts
        assert fn(nat.tile_array([t]),1,out.ctypes.data,nat.DTYPES[dtype],None)==0
        want=oracle.render_chunk_counts(level,mrd,ir,ii,d,dtype=dtype,x0=x0,y0=y0,w=w,h=h)
        if u8: want=oracle.counts_to_u8(want,mrd)
        idx=off+np.arange(h)[:,None]*pitch+np.arange(w)[None,:]
        if not np.array_equal(out[idx],want):
            print("MISMATCH",dict(level=level,ir=ir,ii=ii,d=d,mrd=mrd,x0=x0,y0=y0,w=w,h=h,pitch=pitch,off=off,dtype=dtype,sc=sc,rf=rf,u8=u8),flush=True); sys.exit(1)
        mask=np.ones(n_out,bool); mask[idx]=False
        if not (out[mask]==(77 if u8 else -7)).all():
            print("WROTE OUTSIDE",dict(level=level,ir=ir,ii=ii,d=d,mrd=mrd,x0=x0,y0=y0,w=w,h=h,pitch=pitch,off=off,dtype=dtype,sc=sc,rf=rf,u8=u8),flush=True); sys.exit(1)
        cases+=1; pix+=w*h
    print("soak ok:",cases,"random tiles,",pix,"pixels",flush=True)


def codec_soak():
    import ctypes as C, time
    import numpy as np
    import chunk_codec as codec
    from test_escape_emulated import build_emulator
    L=C.CDLL(build_emulator())
    L.emul_serialize_chunks.argtypes=[C.c_void_p, C.c_size_t, C.c_uint32, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_int, C.c_int, C.c_uint]
    L.emul_set_schedule.argtypes=[C.c_uint64]
    rng=np.random.default_rng(int(sys.argv[1])); t_end=time.time()+float(sys.argv[2]); cases=0; total=0
    while time.time()<t_end:
        n=int(rng.choice([rng.integers(1,200), rng.integers(200,20000), rng.integers(20000,400000)]))
        parts=[]; left=n
        while left>0:
            kind=rng.integers(0,5); seg=int(min(left, rng.integers(1, max(2,n//2+1))))
            if kind==0: p=rng.integers(0,256,seg)
            elif kind==1: p=np.repeat(rng.integers(0,3,seg//2+1), rng.integers(1,6,seg//2+1))[:seg]
            elif kind==2: p=np.full(seg,rng.integers(0,256))
            elif kind==3: p=np.arange(seg)%2
            else: p=np.repeat(rng.integers(0,256,seg//3000+1), rng.integers(1,9000,seg//3000+1))[:seg]
            if len(p)<seg: p=np.concatenate([p,np.full(seg-len(p),9)])
            parts.append(p.astype(np.uint8)); left-=seg
        pl=np.ascontiguousarray(np.concatenate(parts)[:n])
        nb=int(rng.integers(1,4)); spans=int(rng.integers(1,3)); keep=int(rng.integers(0,3)); gy=int(rng.choice([65535,1,2,5]))
        in_off=int(rng.integers(0,16)); out_off=int(rng.integers(0,16))
        L.emul_set_schedule(int(rng.integers(0,2**62)) if rng.random()<0.3 else 0)
        ps=(n+15)//16*16+16
        buf=np.zeros(nb*ps+64,np.uint8); a=(-buf.ctypes.data)%16+in_off; src=buf[a:a+nb*ps]
        chunks=[pl]+[np.ascontiguousarray(rng.permutation(pl)) if rng.random()<0.5 else pl[::-1].copy() for _ in range(nb-1)]
        for k,c in enumerate(chunks): src[k*ps:k*ps+n]=c
        stride=(n+1+255)//256*256
        ob=np.full(nb*stride+64,0xEE,np.uint8); b=(-ob.ctypes.data)%16+out_off; out=ob[b:b+nb*stride]
        info=np.zeros(nb*32,np.uint8)
        assert L.emul_serialize_chunks(src.ctypes.data,n,nb,ps,out.ctypes.data,stride,info.ctypes.data,keep,spans,gy)==0
        for k,c in enumerate(chunks):
            ln=int(info[32*k:32*k+8].view('<u8')[0]); o=out[k*stride:(k+1)*stride]
            if o[:ln].tobytes()!=codec.serialize(c) or not (o[n+1:]==0xEE).all():
                print("MISMATCH",n,nb,spans,keep,gy,in_off,out_off,k,flush=True); sys.exit(1)
        cases+=1; total+=n*nb
    print("codec soak ok:",cases,"batches,",total,"bytes",flush=True)


if __name__ == "__main__":
    escape_soak() if MODE == "escape" else codec_soak()
