// native-als: sbt module a PredictionIO checkout adds next to core/ and data/ (settings mirror the templates' build.sbt,
// examples/scala-parallel-recommendation/blacklist-items/build.sbt:20-25).
name := "apache-predictionio-native-als"

organization := "org.apache.predictionio"

scalaVersion := "2.11.12"

libraryDependencies ++= Seq(
  "org.apache.predictionio" %% "apache-predictionio-core" % "0.14.0" % "provided",
  "org.apache.spark"        %% "spark-mllib"              % "2.4.0"  % "provided")

// the JNI shim is built by `make -C native-als` (needs $JAVA_HOME); both shared objects go on java.library.path
javaOptions += "-Djava.library.path=" + (baseDirectory.value / "lib").getAbsolutePath
