"""Stub: src/datamodules/create_dataset.py imports SimpleITK at module level and sets the global threader (:6); only
vol2slice (pure tensor indexing) is exercised by oracle/make_golden.py, nothing of SimpleITK itself."""


class ProcessObject:
    @staticmethod
    def SetGlobalDefaultThreader(name):
        return None
