#!/usr/bin/env python
"""Developer aid: run one synthetic case through the GPU path and show where it differs from the oracle."""
import os, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import synth_cases, oracle_lib as ol
from alac_b200_loader import load_package
pkg = load_package()
want = sys.argv[1]
for gen in (synth_cases.exotic_cases, synth_cases.frame_length_cases, synth_cases.entropy_edge_cases, synth_cases.hostile_cases):
    for name, ocfg, packets in gen():
        if name != want: continue
        cfg = pkg.ParseMagicCookie(ol.make_cookie(ocfg)); dec = pkg.NewPacketDecoder(cfg, 0)
        pcm, errs = dec.DecodePackets(packets)
        for i, p in enumerate(packets):
            st, ref = ol.decode_packet(ocfg, p)
            if ref is None or pcm[i] is None: continue
            a = np.frombuffer(pcm[i], np.uint8); b = np.frombuffer(ref, np.uint8)
            d = np.nonzero(a != b)[0]
            if len(d):
                fb = ocfg.num_channels * ocfg.bps()
                print(name, 'packet', i, 'len', len(p), 'ndiff', len(d), 'first', d[:24], 'frames', sorted(set((d // fb).tolist()))[:40])
                print(' gpu', a[d[:24]], 'ref', b[d[:24]])
