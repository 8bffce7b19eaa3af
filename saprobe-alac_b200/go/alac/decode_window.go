// Streaming Decoder with a GPU read-ahead window: replaces the per-packet loop of decode.go:127-190.
// NOT COMPILED IN THIS REPO'S IMAGE (no Go toolchain). NewDecoder, Format, Duration, Position and Seek keep the
// reference's code (decode.go:50-124); only the refill below changes: instead of decoding one packet per loop
// iteration (decode.go:159-186) it reads the next `Window` packets from the io.ReadSeeker, decodes them with one
// DecodePackets call and serves Read from the result.
package alac

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
	errs  []error
}

// refill is called by Read when s.buf is drained and s.sampleIdx < len(s.samples).
func (s *Decoder) refill(win *windowState, samples []mp4int.SampleInfo) error {
	idx := s.sampleIdx
	if idx < win.base || idx >= win.base+len(win.ready) {
		hi := min(len(samples), idx+Window)
		packets := make([][]byte, 0, hi-idx)

		var readErr error

		for k := idx; k < hi; k++ {
			packet := make([]byte, samples[k].Size)

			if _, err := s.reader.Seek(int64(samples[k].Offset), io.SeekStart); err != nil {
				readErr = fmt.Errorf("seeking to sample %d at offset %d: %w", k, samples[k].Offset, err)

				break
			}

			if _, err := io.ReadFull(s.reader, packet); err != nil {
				readErr = fmt.Errorf("reading sample %d: %w", k, err)

				break
			}

			packets = append(packets, packet)
		}

		win.ready, win.errs = s.dec.DecodePackets(packets)
		if readErr != nil {
			win.ready = append(win.ready, nil)
			win.errs = append(win.errs, readErr)
		}

		win.base = idx
	}

	if err := win.errs[idx-win.base]; err != nil {
		if win.ready[idx-win.base] == nil && idx-win.base == len(win.ready)-1 && len(win.errs) > 0 {
			return err
		}

		return fmt.Errorf("decoding packet %d: %w", idx, err)
	}

	s.buf = win.ready[idx-win.base]
	s.bufOff = 0
	s.sampleIdx++

	return nil
}
