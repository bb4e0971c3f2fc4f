"""TEST PLUMBING: a zero-copy torch view of a device buffer of int32 words (CUDA array interface), used by the GPU tests that stand
in for NCCL on one GPU (several sessions = several ranks; the bitmap SUM is what comm.cu's ncclAllReduce does between real ranks)."""


class DeviceWords:
    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<i4", "data": (int(ptr), False), "version": 2}
