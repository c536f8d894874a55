import hashlib
import json


def sha256sum(filename, skip_header=0, block_size=65536):
    h = hashlib.sha256()
    with open(filename, "rb") as f:
        for chunk in iter(lambda: f.read(block_size), b""):
            h.update(chunk)
    return h.hexdigest(), None


def hash_dict(a_dict):
    return hashlib.md5(json.dumps(a_dict, sort_keys=True).encode()).hexdigest()
