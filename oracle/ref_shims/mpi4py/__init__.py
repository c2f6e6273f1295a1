class _Comm:
    rank = 0
    size = 1

    def Get_rank(self):
        return 0

    def Get_size(self):
        return 1

    def bcast(self, x, root=0):
        return x

    def gather(self, x, root=0):
        return [x]


class MPI:
    COMM_WORLD = _Comm()
