"""Regenerates dbfile_expected.json: an independent numpy parse of the BLAST DB fixtures (run where
/root/reference exists).  The expectations are what the reference's own reader would report
(GetNumOIDs / GetVolumeLength / GetMaxLength / GetSeqLengthExact)."""
import json
import os
import struct

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/c++/src/algo/blast/unit_tests"
VOLUMES = {
    "ntshort": os.path.join(HERE, "ntshort"),
    "nt.41646578": os.path.join(HERE, "nt.41646578"),
    "seqn": os.path.join(REF, "api/data/seqn"),
}


def parse(prefix):
    b = open(prefix + ".nin", "rb").read()
    o = 0
    version, seqtype = struct.unpack_from(">II", b, o); o += 8
    (n,) = struct.unpack_from(">I", b, o); o += 4; title = b[o:o + n].decode(); o += n
    (n,) = struct.unpack_from(">I", b, o); o += 4 + n
    (nseq,) = struct.unpack_from(">I", b, o); o += 4
    (total,) = struct.unpack_from("<Q", b, o); o += 8
    (maxlen,) = struct.unpack_from(">I", b, o); o += 4
    arr = np.frombuffer(b, dtype=">u4", count=3 * (nseq + 1), offset=o).reshape(3, nseq + 1).astype(np.int64)
    nsq = np.fromfile(prefix + ".nsq", dtype=np.uint8)
    start, end = arr[1][:-1], arr[2][:-1]
    lens = (end - start - 1) * 4 + (nsq[end - 1] & 3)
    assert version == 4 and seqtype == 0 and int(lens.sum()) == total and int(lens.max()) == maxlen
    return {"title": title, "n_seq": int(nseq), "total_bases": int(total), "max_len": int(maxlen),
            "first_offsets": start[:8].tolist(), "first_lengths": lens[:8].tolist(),
            "length_checksum": int((lens * (np.arange(nseq) + 1)).sum())}


if __name__ == "__main__":
    out = {k: parse(v) for k, v in VOLUMES.items() if os.path.exists(v + ".nin")}
    json.dump(out, open(os.path.join(HERE, "dbfile_expected.json"), "w"), indent=1)
    print(json.dumps(out, indent=1))
