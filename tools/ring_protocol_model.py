"""Model of the tensor regime's shared-memory ring protocol (tensor_regime.cu): one TMA producer, TWO MMA issuers on
alternate tiles, mbarrier parity waits, nbuf accumulators drained in order by the epilogue.  Random interleaving of the
agents; TMA loads may LAND OUT OF ORDER (30 % of the completions pick a random in-flight load).  Reports, per
(ring stages, stages per tile, accumulators), whether an issuer ever passes a full-barrier wait on a stage that
does not hold its tile's data -- the race behind the intermittent launch failures of short rings.
With in-order landing (replace the random pick by j = 0) every geometry passes.
    python tools/ring_protocol_model.py"""
import random, sys
def sim(n_ring, per_tile, n_tiles, nbuf, seed, epi_fast=True):
    rng = random.Random(seed)
    full_c = [0]*n_ring; empty_c = [0]*n_ring      # completed phases
    content = [None]*n_ring                         # (tile, ks) landed
    pending_tma = []                                # (stage, tag) in flight
    accf_c = [0]*nbuf; acce_c=[0]*nbuf; acce_arr=[0]*nbuf
    acc_content=[None]*nbuf
    def wait(c, P): return (c & 1) != P
    # producer state
    P = dict(t=0, ks=0, s=0, ph=0)
    # issuers
    I = [dict(it=0, s=0, ph=0, ks=0, state='start', id=i, pending_commit=[]) for i in range(2)]
    # mma completion queue: list of (issuer, kind, idx) completing in order per issuer
    mmaq = []   # (issuer_id, action, arg)
    E = dict(it=0)  # epilogue (single agent standing for 4 warps)
    steps=0
    done_tiles=0
    while done_tiles < n_tiles:
        steps+=1
        if steps>2000000: return "deadlock/too long"
        agents=['prod','tma','i0','i1','mma','epi']
        a=rng.choice(agents)
        if a=='prod' and P['t']<n_tiles:
            s=P['s']
            if wait(empty_c[s], P['ph']^1):
                pending_tma.append((s,(P['t'],P['ks'])))
                if ++P['s'] is None: pass
                P['s']+=1
                if P['s']==n_ring: P['s']=0; P['ph']^=1
                P['ks']+=1
                if P['ks']==per_tile: P['ks']=0; P['t']+=1
        elif a=='tma' and pending_tma:
            j=rng.randrange(len(pending_tma)) if rng.random()<0.3 else 0
            s,tag=pending_tma.pop(j)
            content[s]=tag; full_c[s]+=1
        elif a in('i0','i1'):
            X=I[int(a[1])]
            if X['it']>=n_tiles: continue
            it=X['it']; buf=it%nbuf; par=(it//nbuf)&1
            mine=(it&1)==X['id']
            if X['state']=='start':
                if not mine:
                    X['s']+=per_tile
                    if X['s']>=n_ring: X['s']-=n_ring; X['ph']^=1
                    X['it']+=1; continue
                if wait(acce_c[buf], par^1): X['state']='stages'; X['ks']=0
            elif X['state']=='stages':
                s=X['s']
                if wait(full_c[s], X['ph']):
                    if content[s]!=(it,X['ks']): return f"BAD DATA issuer{X['id']} tile {it} ks {X['ks']} stage {s} has {content[s]} full_c={full_c[s]} ph={X['ph']}"
                    mmaq.append((X['id'],'empty',s))
                    if X['ks']==per_tile-1: mmaq.append((X['id'],'accf',(buf,it)))
                    X['s']+=1
                    if X['s']==n_ring: X['s']=0; X['ph']^=1
                    X['ks']+=1
                    if X['ks']==per_tile: X['state']='start'; X['it']+=1
        elif a=='mma' and mmaq:
            # complete oldest of a random issuer
            ids=[m[0] for m in mmaq]; who=rng.choice(ids)
            k=[i for i,m in enumerate(mmaq) if m[0]==who][0]
            _,act,arg=mmaq.pop(k)
            if act=='empty': empty_c[arg]+=1
            else:
                buf,it=arg; accf_c[buf]+=1; acc_content[buf]=it
        elif a=='epi' and E['it']<n_tiles:
            it=E['it']; buf=it%nbuf; par=(it//nbuf)&1
            if wait(accf_c[buf], par):
                if acc_content[buf]!=it: return f"BAD ACC tile {it} buf has {acc_content[buf]}"
                acce_c[buf]+=1; E['it']+=1; done_tiles+=1
    return "ok"
for n_ring,per_tile in [(7,3),(6,3),(5,3),(4,3),(3,2),(4,2)]:
    res=set()
    for seed in range(300):
        res.add(sim(n_ring,per_tile,40,2,seed))
    print(n_ring,per_tile,res)
print("---- nbuf 4")
for n_ring,per_tile in [(5,2),(7,2),(10,2),(14,2),(4,2),(3,2),(6,3),(7,3),(10,3),(14,3)]:
    res=set()
    for seed in range(200):
        res.add(sim(n_ring,per_tile,60,4,seed))
    print(n_ring,per_tile,4,[r[:60] for r in res])
