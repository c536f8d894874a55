"""Drop the b200 engine into an installed clearml-serving (the reference) without patching it.

    import clearml_serving_b200.integration as b2s
    b2s.register_with_reference()            # adds engine_type "b200"
    b2s.register_with_reference("triton")    # or shadow an existing engine name -- "triton", "xgboost", "lightgbm",
                                             # "sklearn": endpoints registered for it are then served by the b200 engine

After this, `clearml-serving model add --engine b200 ...` validates (endpoints.py:5-8 consults the
registry) and ModelRequestProcessor.process_request builds the engine lazily on the first request
(model_request_processor.py:287-291), exactly like the built-in engines.  See INTEGRATION.md.
"""
from .preprocess_service import B200EngineMixin


def register_with_reference(engine_name="b200"):
    from clearml_serving.serving.preprocess_service import BasePreprocessRequest as RefBase

    @RefBase.register_engine(engine_name, modules=["numpy"])
    class B200PreprocessRequest(B200EngineMixin, RefBase):
        def __init__(self, model_endpoint, task=None):
            RefBase.__init__(self, model_endpoint=model_endpoint, task=task)
            self._b200_setup()

    return B200PreprocessRequest
