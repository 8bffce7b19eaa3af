// LibraryDecoder: many tracks of mixed cookies in one call, sharded over the CUDA devices of the box
// (BASELINE configs[4]). NOT COMPILED IN THIS REPO'S IMAGE (no Go toolchain).
//
// A PacketDecoder holds one cookie (decoder.go:79-87); a music library mixes 16- and 24-bit tracks. The library handle
// checks every track like ParseMagicCookie + NewPacketDecoder, splits the tracks into contiguous ranges balanced by
// compressed bytes over the devices (one submitting host thread per device inside libalacb200, no collective), groups
// them by config inside a device and reads every track's packets in place from the bytes handed over (the M4A file
// image with internal/mp4's sample table, or packed packets).
package alac

/*
#include "alac_b200.h"
*/
import "C"

import (
	"fmt"
	"runtime"
	"unsafe"

	mp4int "github.com/mycophonic/saprobe-alac/internal/mp4"
)

// TrackInput is one track of a library batch.
type TrackInput struct {
	Cookie  []byte               // magic cookie (FindALACTrack's first result)
	Data    []byte               // the bytes the packets live in: file image or packed packets (pin them for async copies)
	Samples []mp4int.SampleInfo  // sample table: packet i = Data[Offset : Offset+Size]
}

// TrackOutput mirrors what NewPacketDecoder + DecodePackets return for the track.
type TrackOutput struct {
	Config PacketConfig
	Err    error    // ErrConfig chain when the track has no decoder (bad cookie, unsupported depth)
	PCM    [][]byte // per packet, nil where Errs[i] != nil
	Errs   []error
	Device int
}

type LibraryDecoder struct {
	h *C.alacb200_library
}

// NewLibraryDecoder binds to the given CUDA devices (e.g. 0..7 on one HGX box).
func NewLibraryDecoder(devices ...int) (*LibraryDecoder, error) {
	if len(devices) == 0 {
		devices = []int{Device}
	}

	devs := make([]C.int, len(devices))
	for i, d := range devices {
		devs[i] = C.int(d)
	}

	var handle *C.alacb200_library
	if rc := C.alacb200_library_create(&devs[0], C.int(len(devs)), &handle); rc != C.ALACB200_OK {
		return nil, fmt.Errorf("%w: %w: %s", ErrConfig, ErrCUDA, C.GoString(C.alacb200_last_error()))
	}

	lib := &LibraryDecoder{h: handle}
	runtime.SetFinalizer(lib, func(l *LibraryDecoder) { l.Close() })

	return lib, nil
}

func (l *LibraryDecoder) Close() {
	if l.h != nil {
		C.alacb200_library_destroy(l.h)
		l.h = nil
	}
}

// DecodeTracks decodes every packet of every track; the result order is the input order.
func (l *LibraryDecoder) DecodeTracks(tracks []TrackInput) ([]TrackOutput, error) {
	count := len(tracks)
	out := make([]TrackOutput, count)

	if count == 0 {
		return out, nil
	}

	var pin runtime.Pinner // Go memory handed to C for the duration of the call
	defer pin.Unpin()

	descs := make([]C.alacb200_track_desc, count)
	type staging struct {
		offsets  []C.uint64_t
		sizes    []C.uint32_t
		outBytes []C.uint32_t
		status   []C.int32_t
		pcm      []byte
		stride   int
	}

	stage := make([]staging, count)

	for idx, track := range tracks {
		n := len(track.Samples)
		st := &stage[idx]
		st.offsets = make([]C.uint64_t, max(n, 1))
		st.sizes = make([]C.uint32_t, max(n, 1))
		st.outBytes = make([]C.uint32_t, max(n, 1))
		st.status = make([]C.int32_t, max(n, 1))

		for k, sample := range track.Samples {
			st.offsets[k] = C.uint64_t(sample.Offset)
			st.sizes[k] = C.uint32_t(sample.Size)
		}

		// size of a decoded packet of THIS track (the C side re-checks the cookie)
		st.stride = 4
		if config, err := ParseMagicCookie(track.Cookie); err == nil {
			if bps := int(C.alacb200_bytes_per_sample(C.uint8_t(config.BitDepth))); bps > 0 {
				st.stride = (int(config.FrameLength)*int(config.NumChannels)*bps + 3) &^ 3
			}
		}

		st.pcm = make([]byte, max(n*st.stride, 1))
		desc := &descs[idx]

		if len(track.Cookie) > 0 {
			pin.Pin(&track.Cookie[0])
			desc.cookie = (*C.uint8_t)(unsafe.Pointer(&track.Cookie[0]))
		}

		desc.cookie_len = C.size_t(len(track.Cookie))

		if len(track.Data) > 0 {
			pin.Pin(&track.Data[0])
			desc.data = (*C.uint8_t)(unsafe.Pointer(&track.Data[0]))
		}

		desc.data_len = C.uint64_t(len(track.Data))
		pin.Pin(&st.offsets[0])
		pin.Pin(&st.sizes[0])
		pin.Pin(&st.outBytes[0])
		pin.Pin(&st.status[0])
		pin.Pin(&st.pcm[0])
		desc.offsets, desc.sizes, desc.n = &st.offsets[0], &st.sizes[0], C.uint32_t(n)
		desc.pcm_out, desc.out_stride = (*C.uint8_t)(unsafe.Pointer(&st.pcm[0])), C.uint64_t(st.stride)
		desc.out_bytes, desc.status = &st.outBytes[0], &st.status[0]
	}

	if rc := C.alacb200_library_decode_tracks(l.h, &descs[0], C.uint32_t(count)); rc != C.ALACB200_OK {
		return out, fmt.Errorf("%w: %w: %s", ErrDecode, ErrCUDA, C.GoString(C.alacb200_last_error()))
	}

	for idx := range tracks {
		desc, st := &descs[idx], &stage[idx]
		res := &out[idx]
		res.Device = int(desc.device)
		res.Config = fromC(desc.config)

		switch {
		case desc.result == C.ALACB200_E_CONFIG:
			res.Err = configStatusError(int32(desc.track_status), res.Config)

			continue
		case desc.result != C.ALACB200_OK:
			res.Err = fmt.Errorf("%w: %w: track %d: rc=%d", ErrDecode, ErrCUDA, idx, int(desc.result))

			continue
		}

		n := len(tracks[idx].Samples)
		res.PCM = make([][]byte, n)
		res.Errs = make([]error, n)

		for k := range n {
			switch st.status[k] {
			case C.ALACB200_ST_OK:
				res.PCM[k] = st.pcm[k*st.stride : k*st.stride+int(st.outBytes[k]) : k*st.stride+int(st.outBytes[k])]
			case C.ALACB200_ST_IO_TRUNCATED:
				res.Errs[k] = fmt.Errorf("reading sample %d: %w", k, errTruncated)
			default:
				res.Errs[k] = statusError(int32(st.status[k]))
			}
		}
	}

	return out, nil
}

var errTruncated = fmt.Errorf("unexpected EOF")

func fromC(c C.alacb200_config) PacketConfig {
	return PacketConfig{
		FrameLength: uint32(c.frame_length), BitDepth: uint8(c.bit_depth), NumChannels: uint8(c.num_channels),
		PB: uint8(c.pb), MB: uint8(c.mb), KB: uint8(c.kb), MaxRun: uint16(c.max_run),
		MaxFrameBytes: uint32(c.max_frame_bytes), AvgBitRate: uint32(c.avg_bit_rate), SampleRate: uint32(c.sample_rate),
	}
}

// configStatusError rebuilds the ErrConfig chains of config.go:61-66 and decoder.go:91-93.
func configStatusError(status int32, config PacketConfig) error {
	switch status & 0xff {
	case C.ALACB200_ST_INVALID_COOKIE:
		return fmt.Errorf("%w: %w", ErrConfig, alacintErrInvalidCookie())
	case C.ALACB200_ST_UNSUPPORTED_VERSION:
		return fmt.Errorf("%w: %w", ErrConfig, alacintErrUnsupportedVersion())
	case C.ALACB200_ST_BIT_DEPTH:
		return fmt.Errorf("%w: %w: %d", ErrConfig, alacintErrBitDepth(), config.BitDepth)
	}

	return fmt.Errorf("%w: %s", ErrConfig, C.GoString(C.alacb200_strerror(C.int32_t(status))))
}
