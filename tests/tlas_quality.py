#!/usr/bin/env python
"""(test infrastructure: uses tests/hostsim) CPU estimate of what a better TLAS topology could buy on the instanced scene C3: wide nodes whose
box a ray crosses (no early termination) for the shipped LBVH topology, a PLOC (windowed agglomerative) topology and a full-sweep
SAH topology, all collapsed to 4-wide nodes by the same greedy rule.  Result (10,000 instances, 3,000 box rays): LBVH 36.4, PLOC r=8
35.7, PLOC r=32 37.1, SAH 33.9 node visits per ray: at most 7 % of the TLAS part, ~3 % of a ray, so the LBVH stays.
usage: python tests/tlas_quality.py [n_instances]"""
import sys, time
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'tests')]
import numpy as np
from raycore_b200 import workloads as W
import hostsim_py as hs, engines

def area(lo,hi):
    d=np.maximum(hi-lo,0); return d[...,0]*d[...,1]+d[...,1]*d[...,2]+d[...,2]*d[...,0]

def lbvh_topology(e):
    nd=e.scene.tlas_nodes2(); n=(len(nd)+1)//2
    # nodes 1..n-1 internal, n..2n-1 leaves; leaf: aabb0_min/aabb0_max = box, child1 = instance
    child=np.zeros((2*n,2),np.int64)
    lo=np.zeros((2*n,3)); hi=np.zeros((2*n,3))
    for k in range(1,2*n):
        r=nd[k-1]
        if k<n: child[k]=(r['child0'],r['child1'])
        else: lo[k]=r['aabb0_min']; hi[k]=r['aabb0_max']
    return n,child,lo,hi

def fit(n,child,lo,hi,root=1):
    # post-order
    order=[]; st=[root]
    while st:
        k=st.pop(); order.append(k)
        if child[k,0]: st+= [child[k,0],child[k,1]]
    for k in reversed(order):
        if child[k,0]:
            a,b=child[k]; lo[k]=np.minimum(lo[a],lo[b]); hi[k]=np.maximum(hi[a],hi[b])

def collapse(child,lo,hi,root):
    """BVH4 by greedy largest-area opening; returns dict node -> list of children (internal ids or leaves)"""
    wide={}
    st=[root]
    A=area(lo,hi)
    while st:
        k=st.pop()
        slots=[child[k,0],child[k,1]]
        while len(slots)<4:
            best=-1;ba=-1
            for i,c in enumerate(slots):
                if child[c,0] and A[c]>ba: ba=A[c];best=i
            if best<0: break
            c=slots[best]; slots[best]=child[c,0]; slots.append(child[c,1])
        wide[k]=slots
        for c in slots:
            if child[c,0]: st.append(c)
    return wide

def count_visits(wide,child,lo,hi,root,rays):
    o=rays['o'].astype(np.float64); d=rays['d'].astype(np.float64)
    inv=1.0/np.where(np.abs(d)>1e-5,d,np.copysign(1e-5,d))
    tot_nodes=0; tot_leaf=0
    for r in range(len(o)):
        oo,ii=o[r],inv[r]
        st=[root]
        while st:
            k=st.pop(); tot_nodes+=1
            for c in wide[k]:
                t0=(lo[c]-oo)*ii; t1=(hi[c]-oo)*ii
                tn=np.minimum(t0,t1).max(); tf=np.maximum(t0,t1).min()
                if max(tn,0.0)<=tf:
                    if child[c,0]: st.append(c)
                    else: tot_leaf+=1
    return tot_nodes/len(o), tot_leaf/len(o)

def sah_cost(wide,child,lo,hi,root):
    A=area(lo,hi); return sum(A[k] for k in wide)/A[root]

def ploc(n,leaf_lo,leaf_hi,radius=16):
    """leaves given in Morton order (index 0..n-1). returns child,lo,hi arrays with ids: leaves n..2n-1 (as lbvh), internal new ids; root id returned"""
    N=2*n
    child=np.zeros((N,2),np.int64); lo=np.zeros((N,3)); hi=np.zeros((N,3))
    lo[n:2*n]=leaf_lo; hi[n:2*n]=leaf_hi
    cl=list(range(n,2*n)); nxt=n-1  # internal ids assigned n-1 down to 1 => root = 1
    while len(cl)>1:
        m=len(cl); L=lo[cl]; H=hi[cl]
        nn=np.zeros(m,np.int64)
        for i in range(m):
            a=max(0,i-radius); b=min(m,i+radius+1)
            ul=np.minimum(L[a:b],L[i]); uh=np.maximum(H[a:b],H[i])
            ar=area(ul,uh); ar[i-a]=np.inf
            nn[i]=a+int(np.argmin(ar))
        out=[]
        for i in range(m):
            j=nn[i]
            if nn[j]==i:
                if i<j:
                    k=nxt; nxt-=1
                    child[k]=(cl[i],cl[j]); lo[k]=np.minimum(L[i],L[j]); hi[k]=np.maximum(H[i],H[j])
                    out.append(k)
            else: out.append(cl[i])
        cl=out
    return child,lo,hi,cl[0]

def sah_topdown(n,leaf_lo,leaf_hi):
    N=2*n
    child=np.zeros((N,2),np.int64); lo=np.zeros((N,3)); hi=np.zeros((N,3))
    lo[n:2*n]=leaf_lo; hi[n:2*n]=leaf_hi
    cen=(leaf_lo+leaf_hi)/2
    nxt=[1]
    def build(ids):
        if len(ids)==1: return n+ids[0]
        k=nxt[0]; nxt[0]+=1
        best=(np.inf,None,None)
        for ax in range(3):
            o=ids[np.argsort(cen[ids,ax],kind='stable')]
            Ll=np.minimum.accumulate(leaf_lo[o],0); Lh=np.maximum.accumulate(leaf_hi[o],0)
            Rl=np.minimum.accumulate(leaf_lo[o][::-1],0)[::-1]; Rh=np.maximum.accumulate(leaf_hi[o][::-1],0)[::-1]
            m=len(o); i=np.arange(1,m)
            cost=area(Ll[:-1],Lh[:-1])*i+area(Rl[1:],Rh[1:])*(m-i)
            j=int(np.argmin(cost))
            if cost[j]<best[0]: best=(cost[j],o[:j+1],o[j+1:])
        a=build(best[1]); b=build(best[2])
        child[k]=(a,b); lo[k]=np.minimum(lo[a],lo[b]); hi[k]=np.maximum(hi[a],hi[b])
        return k
    sys.setrecursionlimit(100000)
    root=build(np.arange(n))
    return child,lo,hi,root

if __name__=='__main__':
    NI=int(sys.argv[1]) if len(sys.argv)>1 else 10000
    verts=W.bumpy_sphere(24)
    xf=W.random_trs(NI,2026,extent=40.0)
    e=engines.HostsimEngine([(verts,None,xf,None)])
    n,child,lo,hi=lbvh_topology(e)
    fit(n,child,lo,hi)
    rays=W.box_rays(3000,7,half=44.0)
    w=collapse(child,lo,hi,1)
    print('LBVH  sah',sah_cost(w,child,lo,hi,1),'visits',count_visits(w,child,lo,hi,1,rays))
    leaf_lo=lo[n:2*n].copy(); leaf_hi=hi[n:2*n].copy()
    for rad in (8,32):
        t=time.time(); c2,l2,h2,root=ploc(n,leaf_lo,leaf_hi,rad)
        w2=collapse(c2,l2,h2,root)
        print('PLOC r',rad,'sah',sah_cost(w2,c2,l2,h2,root),'visits',count_visits(w2,c2,l2,h2,root,rays),'t',time.time()-t)
    t=time.time(); c3,l3,h3,root=sah_topdown(n,leaf_lo,leaf_hi)
    w3=collapse(c3,l3,h3,root)
    print('SAH   sah',sah_cost(w3,c3,l3,h3,root),'visits',count_visits(w3,c3,l3,h3,root,rays),'t',time.time()-t)
