GradientTransformation = object
