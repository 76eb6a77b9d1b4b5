"""Minimal HDF5 (superblock v0, v1 object headers, symbol-table groups, contiguous datasets, v1 attribute
messages) WRITER, used only by the tests to fabricate single-read fast5 files from the committed DAQ
fixtures -- the bundled `data/reads/*.fast5` of the reference cannot travel to the GPU box."""
import struct

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF


def _pad8(b):
    return b + b'\0' * (-len(b) % 8)


def _msg(mtype, data):
    data = _pad8(data)
    return struct.pack('<HHB3x', mtype, len(data), 0) + data


def _dataspace(shape):
    return struct.pack('<BBB5x', 1, len(shape), 0) + b''.join(struct.pack('<Q', d) for d in shape)


def _datatype(dtype):
    dtype = np.dtype(dtype)
    if dtype.kind in 'iu':
        bits0 = 0x08 if dtype.kind == 'i' else 0x00
        return struct.pack('<BBBBI', 0x10 | 0, bits0, 0, 0, dtype.itemsize) + struct.pack('<HH', 0, 8 * dtype.itemsize)
    if dtype.kind == 'f':
        assert dtype.itemsize == 8
        return struct.pack('<BBBBI', 0x10 | 1, 0x20, 0x3f, 0, 8) + struct.pack('<HHBBBBI', 0, 64, 52, 11, 0, 52, 1023)
    if dtype.kind == 'S':
        return struct.pack('<BBBBI', 0x10 | 3, 0, 0, 0, dtype.itemsize)
    raise TypeError(dtype)


def _attribute(name, value):
    if isinstance(value, bytes):
        arr = np.array(value, dtype='S{}'.format(max(1, len(value))))
    else:
        arr = np.array(value, dtype=np.float64)
    nm = name.encode() + b'\0'
    dt, ds = _datatype(arr.dtype), _dataspace(())
    body = struct.pack('<BxHHH', 1, len(nm), len(dt), len(ds)) + _pad8(nm) + _pad8(dt) + _pad8(ds) + arr.tobytes()
    return _msg(0x000C, body)


class H5Writer(object):
    def __init__(self):
        self.buf = bytearray(96)          # superblock (56 bytes) + root symbol table entry (40 bytes)

    def _alloc(self, data):
        self.buf += b'\0' * (-len(self.buf) % 8)
        addr = len(self.buf)
        self.buf += data
        return addr

    def _header(self, messages):
        body = b''.join(messages)
        head = struct.pack('<BxHII4x', 1, len(messages), 1, len(body))
        return self._alloc(head + body)

    def dataset(self, array, attrs=None):
        array = np.ascontiguousarray(array)
        addr = self._alloc(array.tobytes())
        msgs = [_msg(0x0001, _dataspace(array.shape)), _msg(0x0003, _datatype(array.dtype)),
                _msg(0x0008, struct.pack('<BBQQ', 3, 1, addr, array.nbytes))]
        msgs += [_attribute(k, v) for k, v in (attrs or {}).items()]
        return self._header(msgs)

    def group(self, entries, attrs=None):
        names = sorted(entries)
        assert len(names) <= 8, "one symbol-table leaf only"
        heap_data = bytearray(b'\0' * 8)
        offsets = {}
        for n in names:
            offsets[n] = len(heap_data)
            heap_data += _pad8(n.encode() + b'\0')
        data_addr = self._alloc(bytes(heap_data))
        heap_addr = self._alloc(b'HEAP' + struct.pack('<B3xQQQ', 0, len(heap_data), UNDEF, data_addr))
        snod = b'SNOD' + struct.pack('<BxH', 1, len(names))
        for n in names:
            snod += struct.pack('<QQII16x', offsets[n], entries[n], 0, 0)
        snod_addr = self._alloc(snod)
        last_key = offsets[names[-1]] if names else 0
        tree = b'TREE' + struct.pack('<BBHQQ', 0, 0, 1, UNDEF, UNDEF) + struct.pack('<QQQ', 0, snod_addr, last_key)
        tree_addr = self._alloc(tree)
        msgs = [_msg(0x0011, struct.pack('<QQ', tree_addr, heap_addr))]
        msgs += [_attribute(k, v) for k, v in (attrs or {}).items()]
        return self._header(msgs)

    def finish(self, root_header, filename):
        sb = b'\x89HDF\r\n\x1a\n' + struct.pack('<BBBBBBBBHHI', 0, 0, 0, 0, 0, 8, 8, 0, 4, 16, 0)
        sb += struct.pack('<QQQQ', 0, UNDEF, len(self.buf), UNDEF)
        sb += struct.pack('<QQII16x', 0, root_header, 0, 0)
        assert len(sb) == 96
        self.buf[:96] = sb
        with open(filename, 'wb') as fh:
            fh.write(bytes(self.buf))


def write_fast5(filename, daq, offset, rng, digitisation, read_number=1):
    """Single-read fast5 with `Raw/Reads/Read_<n>/Signal` (int16) and `UniqueGlobalKey/channel_id` scaling."""
    w = H5Writer()
    signal = w.dataset(np.asarray(daq, dtype=np.int16))
    read = w.group({'Signal': signal}, attrs={'read_number': float(read_number), 'duration': float(len(daq))})
    reads = w.group({'Read_{}'.format(read_number): read})
    raw = w.group({'Reads': reads})
    channel = w.group({}, attrs={'offset': float(offset), 'range': float(rng), 'digitisation': float(digitisation),
                                 'sampling_rate': 4000.0, 'channel_number': b'1'})
    ugk = w.group({'channel_id': channel})
    root = w.group({'Raw': raw, 'UniqueGlobalKey': ugk})
    w.finish(root, filename)
