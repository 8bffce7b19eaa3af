"""Reader for tests/golden/ffmpeg_fixtures.npz (written by tests/golden/gen_ffmpeg_fixtures.py)."""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def load_fixtures():
    z = np.load(os.path.join(HERE, 'golden', 'ffmpeg_fixtures.npz'))
    index = json.loads(bytes(z['index']).decode())
    out = {}
    for e in index:
        name = e['name']
        blob = bytes(z[name + '/packets'])
        sizes = z[name + '/sizes']
        packets, pos = [], 0
        for s in sizes:
            packets.append(blob[pos:pos + int(s)])
            pos += int(s)
        out[name] = dict(meta=e, cookie=bytes(z[name + '/cookie']), packets=packets)
    return out
