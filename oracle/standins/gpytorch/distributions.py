class MultivariateNormal:
    def __init__(self, mean, covariance_matrix):
        self.mean = mean
        self.covariance_matrix = covariance_matrix
