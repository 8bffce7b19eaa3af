"""Hashing of the synthetic suites (tests/synth_cases.py) for the committed drift pin tests/golden/synth_hashes.json."""
import hashlib
import json
import os
import struct

import numpy as np

import oracle_lib as ol
import synth_cases

HERE = os.path.dirname(os.path.abspath(__file__))
PIN_PATH = os.path.join(HERE, 'golden', 'synth_hashes.json')


def all_cases():
    """-> [(suite/name, cfg, packets)] in a fixed order."""
    out = []
    for suite, fn in (('exotic', synth_cases.exotic_cases), ('framelen', synth_cases.frame_length_cases),
                      ('entropy', synth_cases.entropy_edge_cases), ('hostile', synth_cases.hostile_cases)):
        for name, cfg, packets in fn():
            out.append((f'{suite}/{name}', cfg, packets))
    return out


def packets_digest(cfg, packets):
    h = hashlib.sha256(bytes(ol.make_cookie(cfg)))
    for p in packets:
        h.update(struct.pack('<I', len(p)))
        h.update(bytes(p))
    return h.hexdigest()


def result_digest(status, out_bytes, pcm_rows):
    """status i32 [n], out_bytes u32 [n], pcm_rows [n, >= out_bytes] u8 -> sha256 over (status, byte count, PCM) per packet."""
    h = hashlib.sha256()
    for i in range(len(status)):
        nb = int(out_bytes[i]) if int(status[i]) == 0 else 0
        h.update(struct.pack('<iI', int(status[i]), nb))
        h.update(np.ascontiguousarray(pcm_rows[i, :nb]).tobytes())
    return h.hexdigest()


def oracle_result(cfg, packets):
    packed, offs, sizes = ol.pack(packets)
    out, nb, st = ol.decode_batch(cfg, packed, offs, sizes, nthreads=4)
    return st, nb, out


def compute_with_oracle():
    cases = {}
    npk = 0
    for name, cfg, packets in all_cases():
        st, nb, out = oracle_result(cfg, packets)
        hist = {}
        for s in st:
            hist[str(int(s) & 0xff)] = hist.get(str(int(s) & 0xff), 0) + 1
        cases[name] = dict(packets=len(packets), packets_sha256=packets_digest(cfg, packets),
                           result_sha256=result_digest(st, nb, out), statuses=hist)
        npk += len(packets)
    return dict(what='oracle results of tests/synth_cases.py (status word, byte count, PCM) -- drift pin, see gen_synth_hashes.py',
                packets=npk, cases=cases)


def load_pin():
    with open(PIN_PATH) as f:
        return json.load(f)
