import msgpack, numpy as np
def _ext(code, data):
    if code == 1:
        shape, dtype, buf = msgpack.unpackb(data, raw=True)
        return np.frombuffer(buf, dtype=np.dtype(dtype.decode())).reshape(shape)
    return msgpack.ExtType(code, data)
def msgpack_restore(b): return msgpack.unpackb(b, ext_hook=_ext, raw=False)
