"""Host logic of the multi-stream feeder (compressd_domain_saliencyprediction_b200/feeder.py): the packed-record reader on a pipe fed by
a stand-in producer (no decoder, no GPU)."""
import os
import threading

import numpy as np

import _oracle as o


def _write_records(fd, recs, chunk):
    """what the patched decoder writes (oracle/hm_dump/build_hm_dump.py, HM_PACKED_FD), here in small chunks to exercise partial reads"""
    buf = b""
    for poc, hdr, byts, tok in recs:
        buf += np.array([poc, hdr.shape[0], tok.size], np.int32).tobytes() + hdr.astype(np.int32).tobytes() + byts.astype(np.uint16).tobytes() + tok.astype(np.int16).tobytes()
    with os.fdopen(fd, "wb", buffering=0) as f:
        for i in range(0, len(buf), chunk):
            f.write(buf[i:i + chunk])


def test_packed_record_reader_round_trip():
    import sys
    sys.path.insert(0, o.ROOT)
    from compressd_domain_saliencyprediction_b200.feeder import read_packed_record
    r = np.random.default_rng(0)
    recs = []
    for poc in range(7):
        nctu = int(r.integers(1, 40)); counts = r.integers(0, 30, (nctu, 3))
        hdr = np.zeros((nctu, 4), np.int32); hdr[:, 1:] = counts; hdr[:, 0] = np.concatenate([[0], np.cumsum(counts.sum(1))[:-1]])
        tok = r.integers(-3000, 3000, int(counts.sum())).astype(np.int16)
        recs.append((poc, hdr, r.integers(0, 4096, nctu).astype(np.uint16), tok))
    for chunk in (1 << 20, 7):
        rfd, wfd = os.pipe()
        t = threading.Thread(target=_write_records, args=(wfd, recs, chunk)); t.start()
        got = []
        with os.fdopen(rfd, "rb") as f:
            while True:
                x = read_packed_record(f)
                if x is None:
                    break
                got.append(x)
        t.join()
        assert len(got) == len(recs)
        for a, b in zip(got, recs):
            assert a[0] == b[0] and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2]) and np.array_equal(a[3], b[3])


def test_packed_record_reader_reports_truncation():
    import sys
    import pytest
    sys.path.insert(0, o.ROOT)
    from compressd_domain_saliencyprediction_b200.feeder import read_packed_record
    rfd, wfd = os.pipe()
    with os.fdopen(wfd, "wb") as f:
        f.write(np.array([0, 4, 10], np.int32).tobytes() + b"\0" * 20)        # header promises 4*18 + 20 bytes
    with os.fdopen(rfd, "rb") as f, pytest.raises(IOError):
        read_packed_record(f)
