import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gblastn_b200 import engine as E, setup as S, synth, abi
from oracle import refdriver as R
from tests.test_ambiguity import _queries_over_runs, _restored, GOLD
task = sys.argv[1] if len(sys.argv) > 1 else "megablast"
nin, nsq = os.path.join(GOLD, "seqn.nin"), os.path.join(GOLD, "seqn.nsq")
info, off, ln = E.dbfile_index(nin, nsq)
first, runs = E.dbfile_ambiguity(nin, nsq)
raw = np.fromfile(nsq, dtype=np.uint8)
vol = synth.Volume(packed=np.concatenate([raw, np.zeros(32, np.uint8)]), byte_off=off, seq_len=ln)
qs = _queries_over_runs(vol, first, runs, np.random.default_rng(11), 40, 110)
r = R.search(qs, vol, R.default_config(task, taps=R.TAP_TRACEBACK, prelim_only=0), ambiguity=(first, runs))
s = S.Setup(qs, task=task, db_length=vol.total_bases, db_num_seqs=vol.n_seqs)
V, Q = E.FileVolume(nin, nsq), E.Query(s.batch)
g = E.prelim_search(V, Q)
got, ops = E.traceback_search(V, Q, s.gap_x_dropoff_final(), g["hsps"])
want = r["tb_final"]
print("n", got.shape[0], want.shape[0])
cols = ("query_index", "oid", "context", "q_off", "q_end", "s_off", "s_end", "score", "num_ident")
L = "ACGTRYMKWSBDHVN-"
for i in range(min(got.shape[0], want.shape[0])):
    a = [int(got[c][i]) for c in cols]; b = want[i, :9].tolist()
    if a != b:
        print("HSP", i, "got", a, "want", b)
        qi, oid, ctx = b[0], b[1], b[2]
        q = qs[qi] if ctx % 2 == 0 else np.array([3,2,1,0,5,4,7,6,8,9,13,12,11,10,14,15],np.uint8)[qs[qi][::-1]]
        sres = _restored(vol, oid, first, runs); spl = vol.bases(oid)
        qo, so = min(a[3], b[3]), min(a[5], b[5])
        print(" q   ", "".join(L[x] for x in q[qo:qo+40]))
        print(" sres", "".join(L[x] for x in sres[so:so+40]))
        print(" s2b ", "".join(L[x] for x in spl[so:so+40]))
        print(" runs of oid", runs[first[oid]:first[oid+1]].tolist()[:20])
        go = ops[got["esp_off"][i]:got["esp_off"][i]+got["esp_n"][i]]
        wo = r["tb_ops"][want[i,13]:want[i,13]+want[i,14]]
        print(" ops got", [(int(x["op_type"]), int(x["num"])) for x in go][:8], "want", wo[:8].tolist())
        # the reference's call for this hsp
        calls = r["tb_calls"]
        for c in calls:
            if c[1] == oid and c[2] == ctx: print(" refcall", c.tolist())
        break
