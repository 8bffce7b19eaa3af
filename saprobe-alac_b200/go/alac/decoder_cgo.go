// Package alac: drop-in replacement of saprobe-alac's decoder.go on top of libalacb200.so (CUDA, B200).
//
// NOT COMPILED IN THIS REPO'S IMAGE (no Go toolchain, see DESIGN.md section 1). It is the cgo binding a
// maintainer of github.com/mycophonic/saprobe-alac adds; every other file of the reference package stays as it
// is: config.go (PacketConfig, ParseMagicCookie), errors.go, format.go, internal/mp4 and internal/alac's error
// sentinels. This file replaces decoder.go; decode_window.go replaces the packet loop of decode.go.
//
// Build: CGO_ENABLED=1, libalacb200.so on the linker path, include/alac_b200.h on the include path.
package alac

/*
#cgo LDFLAGS: -lalacb200
#include <stdlib.h>
#include "alac_b200.h"
*/
import "C"

import (
	"fmt"
	"runtime"
	"slices"
	"unsafe"

	alacint "github.com/mycophonic/saprobe-alac/internal/alac"
)

var alacBitDepths = []uint8{16, 20, 24, 32}

// PacketDecoder decodes ALAC packets on one CUDA device. Same contract as the reference type
// (decoder.go:79-87): not goroutine-safe, one decoder per goroutine, several decoders per device are fine.
type PacketDecoder struct {
	config PacketConfig
	format PCMFormat
	h      *C.alacb200_decoder
}

// Device selects the CUDA device NewPacketDecoder binds to (process-wide default 0).
var Device = 0

func toC(config PacketConfig) C.alacb200_config {
	return C.alacb200_config{
		frame_length: C.uint32_t(config.FrameLength), bit_depth: C.uint8_t(config.BitDepth),
		num_channels: C.uint8_t(config.NumChannels), pb: C.uint8_t(config.PB), mb: C.uint8_t(config.MB),
		kb: C.uint8_t(config.KB), max_run: C.uint16_t(config.MaxRun), max_frame_bytes: C.uint32_t(config.MaxFrameBytes),
		avg_bit_rate: C.uint32_t(config.AvgBitRate), sample_rate: C.uint32_t(config.SampleRate),
	}
}

// NewPacketDecoder mirrors decoder.go:90-110.
func NewPacketDecoder(config PacketConfig) (*PacketDecoder, error) {
	if !slices.Contains(alacBitDepths, config.BitDepth) {
		return nil, fmt.Errorf("%w: %w: %d", ErrConfig, alacint.ErrBitDepth, config.BitDepth)
	}

	cfg := toC(config)

	var (
		handle *C.alacb200_decoder
		status C.int32_t
	)

	if rc := C.alacb200_create(&cfg, C.int(Device), &handle, &status); rc != C.ALACB200_OK {
		if rc == C.ALACB200_E_CONFIG {
			return nil, fmt.Errorf("%w: %s", ErrConfig, C.GoString(C.alacb200_strerror(status)))
		}

		return nil, fmt.Errorf("%w: cuda: %s", ErrConfig, C.GoString(C.alacb200_last_error()))
	}

	dec := &PacketDecoder{
		config: config,
		format: PCMFormat{SampleRate: int(config.SampleRate), BitDepth: int(config.BitDepth), Channels: int(config.NumChannels)},
		h:      handle,
	}
	runtime.SetFinalizer(dec, func(d *PacketDecoder) { d.Close() })

	return dec, nil
}

// Close releases the device resources (the reference type has nothing to release).
func (d *PacketDecoder) Close() {
	if d.h != nil {
		C.alacb200_destroy(d.h)
		d.h = nil
	}
}

// Format mirrors decoder.go:112-114.
func (d *PacketDecoder) Format() PCMFormat { return d.format }

// statusError rebuilds the reference's %w chain from a status word so errors.Is(err, ErrDecode) and
// errors.Is(err, alacint.ErrBitstreamOverrun) keep working (decoder.go:144, :156, :172, :180, :184, :189).
func statusError(status int32) error {
	var sentinel error

	switch status & 0xff {
	case C.ALACB200_ST_UNSUPPORTED_ELEMENT:
		sentinel = alacint.ErrUnsupportedElement
	case C.ALACB200_ST_INVALID_HEADER:
		sentinel = alacint.ErrInvalidHeader
	case C.ALACB200_ST_INVALID_SHIFT:
		sentinel = alacint.ErrInvalidShift
	case C.ALACB200_ST_BITSTREAM_OVERRUN:
		sentinel = alacint.ErrBitstreamOverrun
	case C.ALACB200_ST_SAMPLE_OVERRUN:
		sentinel = alacint.ErrSampleOverrun
	default: // ALACB200_ST_REF_PANIC: the pure-Go decoder panics here; the GPU path reports a decode error
		sentinel = ErrMalformedPacket
	}

	err := sentinel

	switch (status >> 12) & 3 {
	case C.ALACB200_ENT_MONO:
		err = fmt.Errorf("entropy decode: %w", err)
	case C.ALACB200_ENT_U:
		err = fmt.Errorf("entropy decode U: %w", err)
	case C.ALACB200_ENT_V:
		err = fmt.Errorf("entropy decode V: %w", err)
	}

	switch (status >> 8) & 0xf {
	case C.ALACB200_CTX_SCE:
		return fmt.Errorf("%w: SCE/LFE: %w", ErrDecode, err)
	case C.ALACB200_CTX_CPE:
		return fmt.Errorf("%w: CPE: %w", ErrDecode, err)
	case C.ALACB200_CTX_DSE:
		return fmt.Errorf("%w: DSE: %w", ErrDecode, err)
	case C.ALACB200_CTX_FIL:
		return fmt.Errorf("%w: FIL: %w", ErrDecode, err)
	}

	return fmt.Errorf("%w: %w", ErrDecode, err)
}

// ErrMalformedPacket is returned where the reference implementation would panic on a hostile packet.
var ErrMalformedPacket = fmt.Errorf("alac: malformed packet")

// ErrCUDA wraps failures of the device or of pinned-memory allocation (there is no CPU fallback).
var ErrCUDA = fmt.Errorf("alac: cuda")

// arena returns the decoder-owned pinned staging buffers (grow-only, alacb200_arena): packets are packed into `in`,
// PCM comes back through `out`. A steady stream of DecodePacket / DecodePackets calls allocates no pinned memory,
// like the reference decoder keeps its scratch (decoder.go:79-87).
func (d *PacketDecoder) arena(inBytes, outBytes int) (in, out []byte, err error) {
	var inPtr, outPtr *C.uint8_t

	if rc := C.alacb200_arena(d.h, C.uint64_t(inBytes), C.uint64_t(outBytes), &inPtr, &outPtr); rc != C.ALACB200_OK {
		return nil, nil, fmt.Errorf("%w: %w: %s", ErrDecode, ErrCUDA, C.GoString(C.alacb200_last_error()))
	}

	if inPtr == nil || outPtr == nil {
		return nil, nil, fmt.Errorf("%w: %w: pinned arena unavailable", ErrDecode, ErrCUDA)
	}

	return unsafe.Slice((*byte)(unsafe.Pointer(inPtr)), max(inBytes, 1)), unsafe.Slice((*byte)(unsafe.Pointer(outPtr)), max(outBytes, 1)), nil
}

// decodeInPlace decodes the packets data[offsets[i] : offsets[i]+sizes[i]] (data: the pinned arena or any other
// host bytes, e.g. a whole file image with its sample table) into the arena's output half.
// status[i] == ALACB200_ST_IO_TRUNCATED marks a packet that lies outside data: a READ error, not a decode error.
func (d *PacketDecoder) decodeInPlace(data []byte, offsets []C.uint64_t, sizes []C.uint32_t, out []byte, stride int,
) (outBytes []C.uint32_t, status []C.int32_t, err error) {
	count := len(sizes)
	outBytes = make([]C.uint32_t, count)
	status = make([]C.int32_t, count)

	rc := C.alacb200_decode_packets(d.h, (*C.uint8_t)(unsafe.Pointer(unsafe.SliceData(data))), C.uint64_t(len(data)),
		&offsets[0], &sizes[0], C.uint32_t(count), (*C.uint8_t)(unsafe.Pointer(unsafe.SliceData(out))), C.uint64_t(stride),
		&outBytes[0], &status[0])
	if rc != C.ALACB200_OK {
		return nil, nil, fmt.Errorf("%w: %w: %s", ErrDecode, ErrCUDA, C.GoString(C.alacb200_last_error()))
	}

	return outBytes, status, nil
}

func (d *PacketDecoder) stride() int { return (int(C.alacb200_max_packet_pcm_bytes(d.h)) + 3) &^ 3 }

// DecodePackets decodes many packets in one GPU call. pcm[i] is a fresh slice of numSamples*channels*bps bytes
// (shorter for a partial last packet) or nil with errs[i] set, i.e. exactly what n calls of DecodePacket return.
func (d *PacketDecoder) DecodePackets(packets [][]byte) ([][]byte, []error) {
	count := len(packets)
	pcm := make([][]byte, count)
	errs := make([]error, count)

	if count == 0 {
		return pcm, errs
	}

	fail := func(err error) ([][]byte, []error) {
		for idx := range errs {
			errs[idx] = err
		}

		return pcm, errs
	}

	// Host packer: the decoder's pinned arena, every packet on a 16-byte boundary.
	offsets := make([]C.uint64_t, count)
	sizes := make([]C.uint32_t, count)
	total := 0

	for idx, packet := range packets {
		offsets[idx] = C.uint64_t(total)
		sizes[idx] = C.uint32_t(len(packet))
		total += (len(packet) + 15) &^ 15
	}

	stride := d.stride()

	in, out, err := d.arena(total, count*stride)
	if err != nil {
		return fail(err)
	}

	for idx, packet := range packets {
		copy(in[int(offsets[idx]):], packet)
	}

	outBytes, status, err := d.decodeInPlace(in[:total], offsets, sizes, out, stride)
	if err != nil {
		return fail(err)
	}

	for idx := range packets {
		if status[idx] != C.ALACB200_ST_OK {
			errs[idx] = statusError(int32(status[idx]))

			continue
		}

		pcm[idx] = slices.Clone(out[idx*stride : idx*stride+int(outBytes[idx])])
	}

	return pcm, errs
}

// DecodePacket mirrors decoder.go:117-128.
func (d *PacketDecoder) DecodePacket(packet []byte) ([]byte, error) {
	pcm, errs := d.DecodePackets([][]byte{packet})

	return pcm[0], errs[0]
}

// sentinels of internal/alac/errors.go:24-33 used by library_cgo.go
func alacintErrInvalidCookie() error      { return alacint.ErrInvalidCookie }
func alacintErrUnsupportedVersion() error { return alacint.ErrUnsupportedVersion }
func alacintErrBitDepth() error           { return alacint.ErrBitDepth }
