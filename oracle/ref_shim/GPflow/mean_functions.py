class Zero:
    def __call__(self, X):
        return 0.0
