import numpy as np, sys
sys.path.insert(0,'.')
from oracle import dsp_oracle as D
from bench import synth_clip_np
F32=np.float32
def cmul(a,b): # float32 complex mul with separate roundings (a,b arrays of complex64 represented as (re,im) float32)
    ar,ai=a; br,bi=b
    return (F32(ar*br)-F32(ai*bi)).astype(F32), (F32(ar*bi)+F32(ai*br)).astype(F32)
def stockham(zr, zi, radices, tw):
    n=zr.shape[-1]; Ns=1
    twr,twi=tw
    for R in radices:
        nb=n//R; tstep=n//(Ns*R)
        j=np.arange(nb); k=j%Ns
        vr=[];vi=[]
        for r in range(R):
            xr=zr[...,j+r*nb]; xi=zi[...,j+r*nb]
            if r>0:
                idx=k*r*tstep
                xr,xi=cmul((xr,xi),(twr[idx],twi[idx]))
            vr.append(xr);vi.append(xi)
        # DFT-R in float32 directly (matrix form using accurate constants) -- emulate butterfly via complex64 ops
        v=np.stack([a+1j*b for a,b in zip(vr,vi)]).astype(np.complex64)
        W=np.exp(-2j*np.pi*np.outer(np.arange(R),np.arange(R))/R).astype(np.complex64)
        out=np.einsum('qr,r...->q...',W,v).astype(np.complex64)
        nzr=np.zeros_like(zr); nzi=np.zeros_like(zi)
        base=(j-k)*R+k
        for q in range(R):
            nzr[...,base+q*Ns]=out[q].real; nzi[...,base+q*Ns]=out[q].imag
        zr,zi=nzr,nzi; Ns*=R
    return zr,zi
x=synth_clip_np(0)[:160*400]
xp=np.pad(x,200,mode='reflect')
T=1+(len(xp)-400)//160
idx=np.arange(400)[None,:]+160*np.arange(T)[:,None]
w=D.hanning(400)
fr=(xp[idx]*w).astype(F32)
truth=np.fft.rfft(fr.astype(np.float64))
pk=np.fft.rfft(fr).astype(np.complex64)
k=np.arange(400); tw=(np.cos(-2*np.pi*k/400).astype(F32), np.sin(-2*np.pi*k/400).astype(F32))
Te=T-(T%2)
zr=fr[0:Te:2].copy(); zi=fr[1:Te:2].copy()
Zr,Zi=stockham(zr,zi,[4,4,5,5],tw)
Z=(Zr+1j*Zi)
kk=np.arange(201); m=(400-kk)%400
Xa=(0.5*(Z[:,kk].real+Z[:,m].real)+1j*0.5*(Z[:,kk].imag-Z[:,m].imag)).astype(np.complex64)
Xb=(0.5*(Z[:,kk].imag+Z[:,m].imag)+1j*0.5*(Z[:,m].real-Z[:,kk].real)).astype(np.complex64)
mine=np.empty((Te,201),np.complex64); mine[0::2]=Xa; mine[1::2]=Xb
tr=truth[:Te]
print("pocketfft  abs err: max %.3e mean %.3e"%(np.abs(pk[:Te]-tr).max(), np.abs(pk[:Te]-tr).mean()))
print("pairpacked abs err: max %.3e mean %.3e"%(np.abs(mine-tr).max(), np.abs(mine-tr).mean()))
print("peak |X|", np.abs(tr).max())
# unpacked complex FFT of a single real frame with same stockham
zr=fr[:Te].copy(); zi=np.zeros_like(zr)
Zr,Zi=stockham(zr,zi,[4,4,5,5],tw); single=(Zr+1j*Zi)[:,:201]
print("single     abs err: max %.3e mean %.3e"%(np.abs(single-tr).max(), np.abs(single-tr).mean()))
print("--- variants (single real frame as complex) ---")
def run(radices, twd=False):
    zr=fr[:Te].copy(); zi=np.zeros_like(zr)
    if twd:
        t=(np.cos(-2*np.pi*k/400), np.sin(-2*np.pi*k/400))
    else: t=tw
    Zr,Zi=stockham(zr,zi,radices,t); s=(Zr+1j*Zi)[:,:201]
    return np.abs(s-tr).max(), np.abs(s-tr).mean()
for rad in ([4,4,5,5],[5,5,4,4],[4,5,4,5],[5,4,5,4],[2,2,2,2,5,5],[20,20] if False else [4,4,5,5]):
    print(rad, "%.3e %.3e"%run(rad))
import scipy.fft
sp=scipy.fft.rfft(fr[:Te])
print("scipy f32  abs err: max %.3e mean %.3e"%(np.abs(sp-tr).max(), np.abs(sp-tr).mean()))
# per-bin-class error: weak bins (|X|<0.2)
wk=np.abs(tr)<0.2
for nm,a in (("scipy",sp),("stockham",single),("pair",mine)):
    e=np.abs(a-tr)
    print(nm,"weak-bin err max %.3e mean %.3e | rel-to-own max %.3e"%(e[wk].max(), e[wk].mean(), (e/np.abs(tr)).max()))
