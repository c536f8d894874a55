class Artifacts(object):
    _hash_block_size = 65536
