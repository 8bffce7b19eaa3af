// Streaming Decoder with a GPU read-ahead window: replaces the per-packet loop of decode.go:127-190.
// NOT COMPILED IN THIS REPO'S IMAGE (no Go toolchain). NewDecoder, Format, Duration, Position and Seek keep the
// reference's code (decode.go:50-124); only the refill below changes: instead of reading and decoding one packet per
// loop iteration (decode.go:159-186) it reads the byte span of the next `Window` packets from the io.ReadSeeker with
// ONE read straight into the decoder's pinned arena (the packets of a track sit back to back in mdat), hands that span
// plus the sample table to one alacb200_decode_packets call -- no per-packet copies, no re-packing -- and serves Read
// from the result.
package alac

/*
#include "alac_b200.h"
*/
import "C"

import (
	"fmt"
	"io"

	mp4int "github.com/mycophonic/saprobe-alac/internal/mp4"
)

// Window is the number of packets decoded per GPU call behind Read (a 4096-frame packet is 43-93 ms of audio).
var Window = 2048

type windowState struct {
	base  int      // sample index of ready[0]
	ready [][]byte // decoded PCM of packets [base, base+len(ready))
	errs  []error  // per packet: nil, a read error (returned as it is) or a decode error (wrapped below)
	read  []bool   // errs[i] is a READ error: decode.go:165-174 returns those without the "decoding packet" prefix
}

// refill is called by Read when s.buf is drained and s.sampleIdx < len(s.samples).
func (s *Decoder) refill(win *windowState, samples []mp4int.SampleInfo) error {
	idx := s.sampleIdx
	if idx < win.base || idx >= win.base+len(win.ready) {
		hi := min(len(samples), idx+Window)
		count := hi - idx
		win.base = idx
		win.ready = make([][]byte, count)
		win.errs = make([]error, count)
		win.read = make([]bool, count)

		// byte span of the window inside the file
		lo, end := samples[idx].Offset, uint64(0)
		for k := idx; k < hi; k++ {
			lo = min(lo, samples[k].Offset)
			end = max(end, samples[k].Offset+uint64(samples[k].Size))
		}

		stride := s.dec.stride()

		in, out, err := s.dec.arena(int(end-lo), count*stride)
		if err != nil {
			return err
		}

		got := 0

		var seekErr, readErr error

		if _, seekErr = s.reader.Seek(int64(lo), io.SeekStart); seekErr == nil {
			got, readErr = io.ReadFull(s.reader, in[:end-lo]) // a short read leaves the tail packets outside in[:got]
		}

		offsets := make([]C.uint64_t, count)
		sizes := make([]C.uint32_t, count)

		for k := range count {
			offsets[k] = C.uint64_t(samples[idx+k].Offset - lo)
			sizes[k] = C.uint32_t(samples[idx+k].Size)
		}

		outBytes, status, err := s.dec.decodeInPlace(in[:got], offsets, sizes, out, stride)
		if err != nil {
			return err
		}

		for k := range count {
			switch {
			case status[k] == C.ALACB200_ST_IO_TRUNCATED: // the reader could not deliver this packet (decode.go:165-174)
				win.read[k] = true

				switch {
				case seekErr != nil:
					win.errs[k] = fmt.Errorf("seeking to sample %d at offset %d: %w", idx+k, samples[idx+k].Offset, seekErr)
				case uint64(offsets[k]) >= uint64(got) && readErr != nil && readErr != io.ErrUnexpectedEOF:
					win.errs[k] = fmt.Errorf("reading sample %d: %w", idx+k, readErr)
				case uint64(offsets[k]) >= uint64(got):
					win.errs[k] = fmt.Errorf("reading sample %d: %w", idx+k, io.EOF) // io.ReadFull: nothing read
				default:
					win.errs[k] = fmt.Errorf("reading sample %d: %w", idx+k, io.ErrUnexpectedEOF)
				}
			case status[k] != C.ALACB200_ST_OK:
				win.errs[k] = statusError(int32(status[k]))
			default:
				win.ready[k] = append([]byte(nil), out[k*stride:k*stride+int(outBytes[k])]...)
			}
		}
	}

	k := idx - win.base
	if err := win.errs[k]; err != nil {
		if win.read[k] {
			return err
		}

		return fmt.Errorf("decoding packet %d: %w", idx, err) // decode.go:181
	}

	s.buf = win.ready[k]
	s.bufOff = 0
	s.sampleIdx++

	return nil
}
