#!/usr/bin/env python
"""WAV/PCM command-line twin of the reference's example decoder (cmd/alac-example-decoder/main.go:40-169) over the
GPU-backed API: reads an M4A/MP4 (file or stdin), writes WAV (default) or raw PCM to stdout, format line on stderr.

    python tools/alac_decode.py [-format wav|pcm] [file.m4a] > out.wav
"""
import argparse
import os
import struct
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def wav_header(fmt, data_len):
    """44-byte PCM WAV header (main.go:118-169). 20-bit audio travels in 24-bit containers."""
    bits = 24 if fmt.BitDepth == 20 else fmt.BitDepth
    block = fmt.Channels * bits // 8
    return (b'RIFF' + struct.pack('<I', 36 + data_len) + b'WAVEfmt ' +
            struct.pack('<IHHIIHH', 16, 1, fmt.Channels, fmt.SampleRate, fmt.SampleRate * block, block, bits) +
            b'data' + struct.pack('<I', data_len))


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split('\n')[0])
    ap.add_argument('-format', '--format', default='wav', choices=['wav', 'pcm'])
    ap.add_argument('-device', '--device', type=int, default=0)
    ap.add_argument('file', nargs='?', help='M4A/MP4 file (stdin when omitted or "-")')
    args = ap.parse_args(argv)
    from alac_b200_loader import load_package
    alac = load_package()
    data = sys.stdin.buffer.read() if args.file in (None, '-') else open(args.file, 'rb').read()
    dec = alac.NewDecoder(data, device=args.device)
    fmt = dec.Format()
    print(f'ALAC: {fmt.SampleRate} Hz, {fmt.BitDepth}-bit, {fmt.Channels} channel(s)', file=sys.stderr)
    pcm = dec.ReadAll()
    out = sys.stdout.buffer
    if args.format == 'wav':
        out.write(wav_header(fmt, len(pcm)))
    out.write(pcm)
    return 0


if __name__ == '__main__':
    sys.exit(main())
