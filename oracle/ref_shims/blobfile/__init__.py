def BlobFile(path, mode="rb"):
    return open(path, mode)


def exists(p):
    import os
    return os.path.exists(p)
