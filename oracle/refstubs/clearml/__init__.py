"""Minimal stand-in for the third-party `clearml` SDK so the reference's hot-path modules import
in a container without a ClearML server. TEST INFRASTRUCTURE ONLY (oracle harness): it is used by
oracle/gen_golden.py to run /root/reference's own dispatch + sklearn engine and record golden
vectors. Nothing in the product imports it.

Names mirror the imports at clearml_serving/serving/model_request_processor.py:16-18 and
clearml_serving/serving/preprocess_service.py:11-13.
"""


class Task(object):
    id = "stub-task"
    artifacts = {}

    @classmethod
    def get_task(cls, *a, **k):
        return cls()

    @classmethod
    def init(cls, *a, **k):
        return cls()

    @classmethod
    def query_tasks(cls, *a, **k):
        return []

    def get_logger(self):
        return _Logger()

    def reload(self):
        pass


class _Logger(object):
    def report_text(self, *a, **k):
        pass


class Model(object):
    _paths = {}

    def __init__(self, model_id=None, **k):
        self.id = model_id

    def get_local_copy(self, *a, **k):
        return Model._paths.get(self.id)


class InputModel(Model):
    pass


class StorageManager(object):
    @staticmethod
    def get_local_copy(remote_url=None, **k):
        return remote_url
